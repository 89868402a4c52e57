"""``pytorch3d.transforms.quaternion_to_matrix`` restated (TEST INFRASTRUCTURE, used by oracle/gen_golden_tool.py only).

pytorch3d is a third-party dependency of the reference (``import pytorch3d.transforms as transform``, TG:52) that is neither in
/root/reference nor installed in this image; the reference's ToolPositioning reward calls exactly one function of it inside a
TorchScript function (TG:1853-1854), so the stand-in has to be real, scriptable source.  Published algorithm (pytorch3d
``transforms/rotation_conversions.py``): the quaternion is read REAL PART FIRST, (r, i, j, k) = q[..., 0..3], two_s = 2 / sum(q * q), and
the nine entries are the usual ones.  (The reference passes Isaac Gym's xyzw quaternions to it; that is the reference's business and
is reproduced by running its code as is.)"""
import torch


def quaternion_to_matrix(quaternions: torch.Tensor) -> torch.Tensor:
    r, i, j, k = torch.unbind(quaternions, -1)
    two_s = 2.0 / (quaternions * quaternions).sum(-1)
    o = torch.stack(
        (
            1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
            two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
            two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j),
        ),
        -1,
    )
    return o.reshape(quaternions.shape[:-1] + (3, 3))
