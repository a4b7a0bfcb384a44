"""Optimisation trajectory bookkeeping (mirror of the reference's ``optimization_state.py``).

Pure host-side record keeping: CPU copies of the field + loss per sampled step, JSON export."""

from __future__ import annotations

import json

import torch


class OptimizationState:
    """Field, loss and step number at one sampled iteration (reference optimization_state.py:6-47)."""

    def __init__(self, deformation_field: torch.Tensor, loss: float, step: int):
        self.deformation_field = deformation_field.detach().cpu()
        self.loss = loss
        self.step = step

    def as_dict(self) -> dict:
        return {"deformation_field": self.deformation_field.tolist(), "loss": self.loss, "step": self.step}


class OptimizationTracker:
    """List of sampled ``OptimizationState``s (reference optimization_state.py:50-144)."""

    def __init__(self, sample_every_n_steps: int, total_steps: int):
        self.checkpoints: list[OptimizationState] = []
        self.sample_every_n_steps = sample_every_n_steps
        self.total_steps = total_steps

    def sample_this_step(self, step: int) -> bool:
        return step % self.sample_every_n_steps == 0 or step == self.total_steps - 1

    def add_checkpoint(self, deformation_field: torch.Tensor, loss: float, step: int) -> None:
        self.checkpoints.append(OptimizationState(deformation_field, loss, step))

    def as_dict(self) -> dict:
        return {
            "optimization_checkpoints": [c.as_dict() for c in self.checkpoints],
            "sample_every_n_steps": self.sample_every_n_steps,
            "total_steps": self.total_steps,
        }

    def to_json(self, filepath: str) -> None:
        with open(filepath, "w") as f:
            json.dump(self.as_dict(), f)
