"""Import shim: the reference scripts do `import utils`; the implementation lives in
faststyle_b200.utils (B200 engine)."""
from faststyle_b200.utils import *  # noqa: F401,F403
