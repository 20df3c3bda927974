# mirrors /root/reference/sihmpc/ihmpccontroller_log.py (module path kept for drop-in imports)
from sihmpc_b200.ihmpccontroller_log import *  # noqa: F401,F403
from sihmpc_b200.ihmpccontroller_log import IHMPCController  # noqa: F401
