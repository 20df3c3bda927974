"""sihmpc_b200 -- B200-native batched closed-loop engine for the sihmpc IHMPC hot path."""
__version__ = "0.1.0"
