"""Top-level alias so the reference's `import lmdb` (build-index.py:5, query-index.py:6)
resolves to clipb200's LMDB-format store when `cli-p_b200/` is on sys.path and no real
py-lmdb is installed (clipb200.lmdb.open defers to a real binding when it finds one)."""
from clipb200.lmdb import (BadValsizeError, Cursor, Environment, Error, MapFullError,  # noqa: F401
                           Transaction, open)
