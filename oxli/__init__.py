"""Drop-in alias: `import oxli; oxli.KmerCountTable(...)` resolves to the
B200-native implementation in oxli_b200 (same class name and methods as the
reference module defined at src/lib.rs:953-957)."""
from oxli_b200 import KmerCountTable, __version__  # noqa: F401

__all__ = ["KmerCountTable"]
