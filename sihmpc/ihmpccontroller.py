# mirrors /root/reference/sihmpc/ihmpccontroller.py (module path kept for drop-in imports)
from sihmpc_b200.ihmpccontroller import *  # noqa: F401,F403
from sihmpc_b200.ihmpccontroller import IHMPCController  # noqa: F401
