"""Default pinhole intrinsics of the 640x480 RGB-D sensor the reference assumes (its ``consts.INTRINSIC``)."""
_FU, _FV = 615.1436, 615.4991          # focal lengths in pixels
_U0, _V0 = 315.3623, 251.5415          # principal point

# 3x4 projection [K | 0]
INTRINSIC = [[_FU, 0.0, _U0, 0.0], [0.0, _FV, _V0, 0.0], [0.0, 0.0, 1.0, 0.0]]
