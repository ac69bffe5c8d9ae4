"""flexynesis_b200 -- B200-native (sm_100a) engine for flexynesis's data-parallel training hot path.

Drop-in classes (same names / constructor signatures / state_dict keys as flexynesis.models):
    DirectPred, supervised_vae, MultiTripletNetwork, GNN, CrossModalPred
Importing this package loads flexynesis_b200/lib/libfxn_b200.so (hand-written CUDA behind a C ABI, see
include/flexynesis_b200.h); the import fails loudly when the library has not been built.
"""
from . import _lib  # noqa: F401  (loads the shared library or raises)
from .models import CrossModalPred, DirectPred, GNN, MultiTripletNetwork, supervised_vae  # noqa: F401
from .data import DeviceBatcher, DeviceTripletBatcher, SyntheticMultiOmicDataset  # noqa: F401
from . import fit, parallel, trials  # noqa: F401,E402

__all__ = ["CrossModalPred", "DirectPred", "GNN", "MultiTripletNetwork", "supervised_vae", "SyntheticMultiOmicDataset", "DeviceBatcher", "DeviceTripletBatcher"]
__version__ = "0.1.0"
