"""cheetah_b200 -- B200-native backend for Cheetah's ``Segment.track(ParticleBeam)`` hot path.

The public names mirror ``cheetah`` (desy-ml/cheetah) for the elements, beams and
species that lie on the path; the arithmetic runs in hand-written sm_100a
CUDA kernels behind the C ABI declared in ``include/cheetah_b200.h``.
"""

from .beam import Beam, ParameterBeam, ParticleBeam  # noqa: F401
from .elements import (  # noqa: F401
    BPM,
    Aperture,
    Cavity,
    CombinedCorrector,
    CustomTransferMap,
    Dipole,
    Drift,
    Element,
    HorizontalCorrector,
    Marker,
    PhysicsWarning,
    Quadrupole,
    RBend,
    Screen,
    Segment,
    Sextupole,
    Solenoid,
    SpaceChargeKick,
    Superimposed,
    TransverseDeflectingCavity,
    Undulator,
    VerticalCorrector,
)
from .graphs import GraphedTrack  # noqa: F401
from .space_charge import cloud_in_cell_charge_deposition  # noqa: F401
from .species import Species  # noqa: F401
from .tracking import BeamMoments, first_order_transfer_map, track, track_moments  # noqa: F401

__version__ = "0.1.0"
