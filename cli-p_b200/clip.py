"""Top-level alias so the reference's `import clip` (build-index.py:7,
query-index.py:8) resolves to the B200 implementation when `cli-p_b200/` is on
sys.path."""
from clipb200.clip import available_models, load, tokenize  # noqa: F401
