from .core import *  # noqa: F401,F403
from .core import (acos_safe, ensure_batch_dim, normalize, uniform, uniform_unit_vector, inner_product,
                   homogenize, dehomogenize, transform_coords, grid_to_coords)  # noqa: F401
from . import quaternion  # noqa: F401
from .rigid import *  # noqa: F401,F403
from .rigid import (intrinsic_to_3x4, matrix_3x3_to_4x4, rotation_to_4x4, translation_to_4x4, decompose,
                    inverse_transform, translate_matrix, scale_matrix, extrinsic_to_position,
                    random_translation, to_extrinsic_matrix, extrinsic_to_quat)  # noqa: F401
from .batchview import bv2b, b2bv, bvmm, vcat, vsplit  # noqa: F401
from . import orientation  # noqa: F401
