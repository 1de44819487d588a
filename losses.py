"""Import shim: the reference scripts do `import losses`; the implementation lives in
faststyle_b200.losses (B200 engine)."""
from faststyle_b200.losses import *  # noqa: F401,F403
