"""B200-native Decision/Planning hot path.  Python side: ABI mirrors, the
synthetic scene generator and a thin ctypes binding of libdmpp_b200.so (the product is the
C-ABI library in csrc/; this package only loads it)."""
from . import abi, scenes  # noqa: F401
