"""Top-level alias so the reference's `import faiss` (build-index.py:8,
query-index.py:9) resolves to the B200 implementation when `cli-p_b200/` is on
sys.path."""
from clipb200.faiss import *  # noqa: F401,F403
from clipb200.faiss import IndexFlatIP, IndexIVFFlat, METRIC_INNER_PRODUCT, METRIC_L2, read_index, write_index  # noqa: F401
