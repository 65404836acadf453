"""B200-native PointNet++ set-abstraction hot path (drop-in for the reference's models/*.py)."""
from . import _lib, ops, losses, sa, models, synthetic, dp, graph, optim, trunk, ssg_msg, pointnet, data  # noqa: F401
from .sa import (PointNetSetAbstraction, index_points, square_distance, query_ball_point,  # noqa: F401
                 set_default_precision, get_default_precision)
from .models import (DIRS_8, PointNetPP8Dir, PointNetPPVonMises, PointNetPPMvM, PointNetPPXYZ,  # noqa: F401
                     PointNetPP, PointNetPPXYZ_Schedmit, PointNetPPFwd, mvm_density_on_grid)
from .losses import kl_von_mises, kl_von_mises_clamped, match_loss, kl_loss_per_sample_from_logits  # noqa: F401
from .ops import farthest_point_sample, ball_query  # noqa: F401
from .ssg_msg import (SimpleSetAbstraction, SimpleSetAbstractionGroupAll, PointNetPlusPlusCls,  # noqa: F401
                      PointNetSetAbstractionMsg)

from .pointnet import STN3d, STNkd, PointNetEncoder, PointNet  # noqa: F401
from .graph import GraphedTrainStep  # noqa: F401

__version__ = "0.1.0"
