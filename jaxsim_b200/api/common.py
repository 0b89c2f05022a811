"""``VelRepr`` and the 6D representation conversions at the step boundary.

Mirrors ``src/jaxsim/api/common.py:39-47`` (``VelRepr``) and ``:100-222``
(``inertial_to_other_representation`` / ``other_representation_to_inertial``).  These run
as small batched torch ops on the device: they are only needed when the caller passes
``link_forces`` / base velocities in a non-inertial representation; the state
itself is always stored inertial-fixed (``api/data.py:36-39``).
"""

from __future__ import annotations

import enum

import torch


class VelRepr(enum.IntEnum):
    """``api/common.py:39-47``."""

    Body = enum.auto()
    Mixed = enum.auto()
    Inertial = enum.auto()


def _wedge(v: torch.Tensor) -> torch.Tensor:
    x, y, z = v[..., 0], v[..., 1], v[..., 2]
    o = torch.zeros_like(x)
    return torch.stack(
        [torch.stack([o, -z, y], -1), torch.stack([z, o, -x], -1), torch.stack([-y, x, o], -1)], -2
    )


def adjoint_from_transform(H: torch.Tensor, inverse: bool = False) -> torch.Tensor:
    """``Adjoint.from_transform`` (``math/adjoint.py:46-107``), batched."""
    R = H[..., 0:3, 0:3]
    p = H[..., 0:3, 3]
    X = torch.zeros(H.shape[:-2] + (6, 6), dtype=H.dtype, device=H.device)
    if not inverse:
        X[..., 0:3, 0:3] = R
        X[..., 0:3, 3:6] = _wedge(p) @ R
        X[..., 3:6, 3:6] = R
    else:
        Rt = R.transpose(-1, -2)
        X[..., 0:3, 0:3] = Rt
        X[..., 0:3, 3:6] = -Rt @ _wedge(p)
        X[..., 3:6, 3:6] = Rt
    return X


def other_representation_to_inertial(
    array: torch.Tensor, other_representation: VelRepr, transform: torch.Tensor, *, is_force: bool
) -> torch.Tensor:
    """``api/common.py:160-222``."""
    if other_representation == VelRepr.Inertial:
        return array
    W_H_O = transform
    if other_representation == VelRepr.Mixed:
        W_H_O = W_H_O.clone()
        W_H_O[..., 0:3, 0:3] = torch.eye(3, dtype=W_H_O.dtype, device=W_H_O.device)
    elif other_representation != VelRepr.Body:
        raise ValueError(other_representation)
    if not is_force:
        X = adjoint_from_transform(W_H_O)
    else:
        X = adjoint_from_transform(W_H_O, inverse=True).transpose(-1, -2)
    return torch.einsum("...ij,...j->...i", X, array)


def inertial_to_other_representation(
    array: torch.Tensor, other_representation: VelRepr, transform: torch.Tensor, *, is_force: bool
) -> torch.Tensor:
    """``api/common.py:100-158``."""
    if other_representation == VelRepr.Inertial:
        return array
    W_H_O = transform
    if other_representation == VelRepr.Mixed:
        W_H_O = W_H_O.clone()
        W_H_O[..., 0:3, 0:3] = torch.eye(3, dtype=W_H_O.dtype, device=W_H_O.device)
    elif other_representation != VelRepr.Body:
        raise ValueError(other_representation)
    if not is_force:
        X = adjoint_from_transform(W_H_O, inverse=True)
    else:
        X = adjoint_from_transform(W_H_O).transpose(-1, -2)
    return torch.einsum("...ij,...j->...i", X, array)
