"""README quick-start DCN on a generated Parquet dataset (BASELINE config 1)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hugectr  # noqa: E402
from hugectr.tools import DataGenerator, DataGeneratorParams  # noqa: E402
from hugectr_b200.models.legacy import build_dcn  # noqa: E402

slots = [10000] * 26
d = "./dcn_parquet_data"
gp = DataGeneratorParams(format=hugectr.DataReaderType_t.Parquet, label_dim=1, dense_dim=13, num_slot=26,
                         i64_input_key=True, source=f"{d}/file_list.txt", eval_source=f"{d}/file_list_test.txt",
                         slot_size_array=slots, dist_type=hugectr.Distribution_t.PowerLaw,
                         power_law_type=hugectr.PowerLaw_t.Short, num_files=4, eval_num_files=1,
                         num_samples_per_file=16384)
if not os.path.exists(gp.source):
    DataGenerator(gp).generate()
model = build_dcn(batchsize=1024, source=gp.source, eval_source=gp.eval_source, slot_sizes=slots)
model.compile()
model.summary()
model.fit(max_iter=5120, display=200, eval_interval=1000, snapshot=5000, snapshot_prefix="dcn")
