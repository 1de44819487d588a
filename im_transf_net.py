"""Import shim: the reference scripts do `import im_transf_net`; the implementation lives in
faststyle_b200.im_transf_net (B200 engine)."""
from faststyle_b200.im_transf_net import *  # noqa: F401,F403
