"""tf1_shim: names Util/Loss.py imports at module load (its functions are never executed on the hot path)."""
from tensorflow import zeros_like, ones_like  # noqa: F401
