# Drop-in import surface of the third-party `opom` package the reference examples import
# (/root/reference/example_distillation.py:11: from opom import OPOM, TransferFunction).
from sihmpc_b200.opom import OPOM, TransferFunction  # noqa: F401
