"""Drop-in for /root/reference/code/rendering/utils.py: small tensor helpers around the kernels
(texture symmetry / pole / wrap handling, quaternion products).  Same names and semantics."""
import torch
import torch.nn.functional as F


def grid_sample_bilinear(input, grid):
    """Bilinear lookup with corner-aligned coordinates (reference :6-12, modern-PyTorch branch)."""
    return F.grid_sample(input, grid, mode='bilinear', align_corners=True)


def symmetrize_texture(x):
    """Half-width map -> full width by even reflection about the x axis (reference :15-18):
    [ mirrored right half | x | mirrored left half ]."""
    w = x.shape[3]
    m = x.flip(3)
    return torch.cat((m[..., w // 2:], x, m[..., :w // 2]), dim=3)


def adjust_poles(tex):
    """Rows 0 and -1 map to the sphere poles: replace each by its mean (reference :21-26)."""
    out = tex.clone()
    out[:, :, 0] = tex[:, :, 0].mean(dim=2, keepdim=True)
    out[:, :, -1] = tex[:, :, -1].mean(dim=2, keepdim=True)
    return out


def circpad(x, amount=1):
    """Wrap-around padding along x only (reference :29-33).  Written as a concatenation: unlike
    F.pad(mode='circular') it keeps a channels-last tensor channels-last, which the conv kernels read directly."""
    return torch.cat((x[..., -amount:], x, x[..., :amount]), dim=3)


def qrot(q, v):
    """Rotate vectors v [B,N,3] by unit quaternions q [B,4] (w,x,y,z) (reference :36-46)."""
    if q.shape[-1] != 4 or v.shape[-1] != 3:
        raise ValueError("qrot expects q [...,4] and v [...,3]")
    w = q[:, None, :1]
    u = q[:, None, 1:].expand(-1, v.shape[1], -1)
    t = torch.cross(u, v, dim=2)
    return v + 2 * (w * t + torch.cross(u, t, dim=2))


def qmul(q, r):
    """Hamilton product q*r of quaternions stored (w,x,y,z) (reference :48-63)."""
    if q.shape[-1] != 4 or r.shape[-1] != 4:
        raise ValueError("qmul expects quaternions on the last axis")
    shape = q.shape
    q, r = q.reshape(-1, 4), r.reshape(-1, 4)
    qw, qx, qy, qz = q.unbind(1)
    rw, rx, ry, rz = r.unbind(1)
    out = torch.stack((qw * rw - qx * rx - qy * ry - qz * rz,
                       qw * rx + qx * rw + qy * rz - qz * ry,
                       qw * ry - qx * rz + qy * rw + qz * rx,
                       qw * rz + qx * ry - qy * rx + qz * rw), dim=1)
    return out.view(shape)
