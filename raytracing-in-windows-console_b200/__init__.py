"""rtc_b200 -- Python plumbing over the C-ABI in include/rtc.h (placeholder until csrc builds)."""
from . import _types, scenes  # noqa: F401
