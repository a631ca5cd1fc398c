"""Isolated embedding fwd/bwd timing on the DLRM-DCNv2 lookup shapes (1 GPU)."""
import sys, torch
sys.path.insert(0, ".")
from hugectr_b200.embedding.collection import EmbeddingCollection, EmbeddingCollectionConfig, EmbeddingTableConfig
from hugectr_b200.models.dlrm import CRITEO_TB_MULTI_HOT as H, CRITEO_TB_TABLE_SIZES as T
from hugectr_b200.parallel.comm import Comm
from hugectr_b200.solver import CreateOptimizer
from hugectr_b200.enums import Optimizer_t
from hugectr_b200.data.batch import power_law_keys
small = "--small" in sys.argv
sizes = [min(t, 100000) for t in T] if small else T
dev = torch.device("cuda")
cfg = EmbeddingCollectionConfig()
ts = [EmbeddingTableConfig(str(i), sizes[i], 128) for i in range(26)]
cfg.embedding_lookup(ts, [f"d{i}" for i in range(26)], "emb", ["sum"] * 26)
b = 6912
e = EmbeddingCollection(cfg, b, {f"d{i}": H[i] for i in range(26)}, dev, torch.bfloat16, Comm(dev),
                        CreateOptimizer(Optimizer_t.AdaGrad, epsilon=1e-8), state_dtype=torch.bfloat16 if not small else torch.float32)
g = torch.Generator().manual_seed(0)
keys = torch.cat([power_law_keys(b * H[i], sizes[i], 1.1, g).int() for i in range(26)]).cuda()
e.set_keys(keys); e.top_grad["emb"].normal_(0, 0.01)
lr = torch.tensor([0.004], device=dev); st = torch.ones(1, dtype=torch.int32, device=dev)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a, c = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(n): fn()
    c.record(); torch.cuda.synchronize(); return a.elapsed_time(c) / n * 1000
print("fwd us", t(e.forward))
print("bwd index us", t(lambda: (e.backward_index(), e.backward(lr, st))) , "(index+reduce+update)")
