"""Frames rigidly attached to links: the subset of ``jaxsim.api.frame`` (``src/jaxsim/api/frame.py``) the weld
constraints of the step need.  A frame index counts after the links: ``index = number_of_links + position``
(``api/frame.py:21-57``)."""

from __future__ import annotations

import numpy as np
import torch


def _position(model, frame_index: int) -> int:
    nL, nF = model.number_of_links(), model.kin_dyn_parameters.number_of_frames()
    k = int(frame_index) - nL
    if not 0 <= k < nF:
        raise ValueError(f"frame index {frame_index} outside [{nL}, {nL + nF})")  # api/frame.py:35-44
    return k


def name_to_idx(model, *, frame_name: str) -> int:
    """``js.frame.name_to_idx`` (``api/frame.py:60-85``)."""
    names = model.kin_dyn_parameters.frame_parameters.name
    if frame_name not in names:
        raise ValueError(f"frame '{frame_name}' not found in the model")
    return model.number_of_links() + names.index(frame_name)


def idx_to_name(model, *, frame_index: int) -> str:
    """``js.frame.idx_to_name`` (``api/frame.py:88-109``)."""
    return model.kin_dyn_parameters.frame_parameters.name[_position(model, frame_index)]


def idx_of_parent_link(model, *, frame_index: int) -> int:
    """``js.frame.idx_of_parent_link`` (``api/frame.py:21-57``)."""
    return int(model.kin_dyn_parameters.frame_parameters.body[_position(model, frame_index)])


def transform(model, data, *, frame_index: int) -> torch.Tensor:
    """``js.frame.transform`` (``api/frame.py:162-194``): ``W_H_F = W_H_L @ L_H_F``, batched like ``data``."""
    k = _position(model, frame_index)
    fp = model.kin_dyn_parameters.frame_parameters
    W_H_L = data.link_transforms[..., int(fp.body[k]), :, :]
    L_H_F = torch.as_tensor(np.asarray(fp.transform)[k], dtype=W_H_L.dtype, device=W_H_L.device)
    return W_H_L @ L_H_F
