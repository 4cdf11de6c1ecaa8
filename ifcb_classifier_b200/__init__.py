"""ifcb_classifier_b200 -- B200-native hot path of WHOIGit/ifcb_classifier.

Python/PyTorch host code over hand-written sm_100a CUDA behind a C ABI
(include/ifcb_b200.h).  PyTorch provides device memory, streams and
torch.distributed only; every kernel on the path is in csrc/.
"""
__version__ = '0.1.0'
