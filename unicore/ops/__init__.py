"""``unicore.ops`` - the op layer used by modules/optimizers; implemented in ``unicore_b200.ops``."""
from unicore_b200.ops import *  # noqa: F401,F403
from unicore_b200.ops import HAS_CUDA_EXT, USE_NATIVE, native, use_native  # noqa: F401
