"""seistorch_b200 -- B200-native (sm_100a) implementation of Seistorch's wave-propagation
hot path behind the reference's own Python plug-in surface.

Public surface (mirrors the reference package):
    build_model, WaveRNN, WaveCell, WaveSource, WaveProbe, TensorList, Loss / L2 / L1 / CosineSimilarity / Envelope,
    equations2d.<eq>._time_step, equations3d.acoustic._time_step, checkpoint
"""
from .type import TensorList  # noqa: F401
from .source import WaveSource  # noqa: F401
from .probe import WaveProbe, WaveIntensityProbe  # noqa: F401
from .cell import WaveCell  # noqa: F401
from .rnn import WaveRNN  # noqa: F401
from .loss import Loss, L2, L1, SML1, Crosscorrelation, Integration, CosineSimilarity, NormalizedIntegrationMethod, Wasserstein1d, Traveltime, Envelope  # noqa: F401
from .model import build_model, model_from_case  # noqa: F401
from .engine import release_buffers  # noqa: F401

__version__ = "0.1.0"
