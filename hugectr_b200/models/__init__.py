"""Model zoo built through the public API (the reference's samples/)."""
from .dlrm import CRITEO_TB_TABLE_SIZES, CRITEO_TB_MULTI_HOT, build_dlrm_dcnv2, build_dlrm  # noqa: F401
from .legacy import build_dcn, build_deepfm, build_wdl  # noqa: F401,E402
from .zoo import (build_bst, build_criteo_dnn, build_din, build_dlrm_ftrl, build_mmoe, build_ncf,  # noqa: F401,E402
                  build_shared_bottom)
