"""Model zoo built through the public API (the reference's samples/)."""
from .dlrm import CRITEO_TB_TABLE_SIZES, CRITEO_TB_MULTI_HOT, build_dlrm_dcnv2, build_dlrm  # noqa: F401
