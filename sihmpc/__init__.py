# Drop-in import surface of the reference package (/root/reference/sihmpc/__init__.py:1):
#   from sihmpc import IHMPCController        -> the WEIGHTED controller
from sihmpc_b200.ihmpccontroller import *  # noqa: F401,F403
