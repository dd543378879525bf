"""KD training module: the hot-path slice of the reference's pl_module.py (NFModel) without Lightning.

Kept: ``_get_kd_indices`` (pl_module.py:81-110), ``load_checkpoint`` (:112-129), ``create_model`` (:131-156),
``forward`` (:198-255), ``loss`` (:257-320), ``generate`` (:322-346), ``configure_optimizers`` (:348-363),
``training_step`` (:365-382) — same config keys, same batch formats, same returned dict. Everything else in the
reference module (validation, FID/KS logging, data hooks) is out of scope.
"""
from __future__ import annotations

import os
import typing as tp

import torch
import torch.nn as nn

from . import functional as Fn
from .models import (FlowStep, SqueezeLayer, create_glow_model, gaussian_sample,  # noqa: F401
                     inherit_permutation_matrix)

TABULAR = ("bsds300", "gas", "hepmass", "miniboone", "power")


class NFModel(nn.Module):
    def __init__(self, config: tp.Dict[str, tp.Any]):
        super().__init__()
        self.params = config
        self.nll_weight = config["loss"]["nll"]["weight"]
        self.kd_weight = config["loss"]["kd"]["weight"]
        self.perceptual_weight = config["loss"]["perceptual"]["weight"]
        self.teacher = self.create_model("teacher")
        self.student = self.create_model("student")
        if self.teacher is not None:
            for p in self.teacher.parameters():
                p.requires_grad_(False)
        name = config["loss"]["perceptual"].get("name", "l1")
        if self.perceptual_weight > 0 and name != "l1":
            raise NameError("only the l1 perceptual loss is on the hot path (vgg is out of scope)")
        if config["loss"]["kd"].get("name", "mse") != "mse":
            raise NameError("Unknown KD loss name")
        self.student_kd_indices, self.teacher_kd_indices = self._get_kd_indices()
        if config.get("inherit_p", False):     # pl_module.py:64-76
            assert not config["teacher"].get("is_1d", False), "Teacher model must be 3-dimensional"
            assert not config["student"].get("is_1d", False), "Student model must be 3-dimensional"
            inherit_permutation_matrix(self.student, self.teacher, self.student_kd_indices,
                                       self.teacher_kd_indices)
        self.logged: tp.Dict[str, torch.Tensor] = {}
        self.concurrent_teacher = os.environ.get("NFK_CONCURRENT_TEACHER", "1") != "0"
        self._side = None

    @property
    def device(self):
        return next(self.student.parameters()).device

    def log(self, name, value, **kwargs):  # Lightning hook stand-in
        self.logged[name] = value.detach()

    # ---- pl_module.py:81-110
    def _get_kd_indices(self):
        if self.kd_weight + self.perceptual_weight == 0:
            return [], []
        is_1d = self.params["student"]["is_1d"]
        mult = 2
        s = [i for i, l in enumerate(self.student.flow.layers)
             if isinstance(l, SqueezeLayer) or (is_1d and (i + 1) % mult == 0)
             or i + 1 == len(self.student.flow.layers)]
        t = [i for i, l in enumerate(self.teacher.flow.layers)
             if isinstance(l, SqueezeLayer) or (is_1d and (i + 1) % (2 * mult) == 0)
             or i + 1 == len(self.teacher.flow.layers)]
        return s, t

    # ---- pl_module.py:112-129
    def load_checkpoint(self, model, checkpoint_path) -> nn.Module:
        state = torch.load(checkpoint_path, map_location="cuda" if torch.cuda.is_available() else "cpu")
        if "state_dict" in state:
            state = {".".join(k.split(".")[1:]): v for k, v in state["state_dict"].items()
                     if k.startswith("student.")}
        model.load_state_dict(state)
        return model

    # ---- pl_module.py:131-156
    def create_model(self, model_name) -> tp.Optional[nn.Module]:
        if model_name == "teacher" and self.kd_weight + self.perceptual_weight == 0:
            return None
        cfg = dict(self.params[model_name])
        ckpt = cfg.pop("checkpoint", None)
        # the reference reads the architecture from the top-level config (pl_module.py:140); a per-model key is also
        # accepted (bench workloads mix nothing, but a MAF teacher with its own key stays expressible)
        arch = cfg.pop("architecture", self.params.get("architecture", "glow"))
        if arch == "maf":   # extension: the reference names MAF (README.md:7) but only ever builds "glow"
            from .models.maf import create_maf_model
            model = create_maf_model(cfg)
        elif arch != "glow":
            raise NameError(f"Unknown architecture: {arch}")
        else:
            model = create_glow_model(cfg)
        if ckpt:
            model = self.load_checkpoint(model, ckpt)
        return model

    # ---- pl_module.py:198-255
    def forward(self, batch, _defer_objective=False):
        """pl_module.py:198-255. `_defer_objective` (used by training_step only): when the student's prior is a fixed
        row, skip the model's own prior / bits-per-dim reduction and hand logdet + last latent to loss(), whose fused
        kernel computes the objective together with the KD terms; "student_nll" is then None and "student_objective"
        = (logdet, (mean_row, logs_row), scale). Called without the flag the dict is the reference's."""
        name = self.params["data"]["name"]
        if name in TABULAR:
            x, y, weights = batch[0], None, None
        elif "drop_weights" not in self.params["data"]:
            x, y = batch
            weights = None
        else:
            x, y, weights = batch
        cond = y if self.params["student"]["y_condition"] else None
        teacher_z = None
        objective = None
        two_streams = (self.kd_weight > 0 and self.concurrent_teacher and x.is_cuda
                       and not self.params["student"]["is_1d"] and not self.params["teacher"]["is_1d"])
        if x.dtype == torch.uint8 and not two_streams:
            # raw pixels on the sequential path: preprocess once (data/src/utils.py:7-18) so that student and teacher
            # share one fp32 batch and the noise accumulates in it like in the reference
            from . import ops
            xf = torch.empty(x.shape, device=x.device, dtype=torch.float32)
            ops.dequant_squeeze(x.contiguous(), None, xf, None)
            x = xf
        defer = _defer_objective and hasattr(self.student, "deferred_objective") \
            and self.student._prior_rows() is not None
        if two_streams:
            student_z, student_nll, teacher_z = self._forward_two_streams(x, cond, defer)
            if defer:
                objective, student_nll = student_nll, None
        else:
            if defer:
                if self.params["student"]["is_1d"]:
                    ld0, sq0 = torch.zeros(x.shape[0], device=x.device, dtype=torch.float32), None
                else:
                    x, ld0, sq0 = self._dequantize(x)
                student_z, ld, rows, scale = self.student.deferred_objective(x, ld0, cond, _sq0=sq0)
                objective, student_nll = (ld, rows, scale), None
            else:
                student_z, student_nll, _ = self.student(x, cond)
            if self.kd_weight > 0:
                with torch.no_grad():
                    teacher_z, _, _ = self.teacher(x, cond)   # x already carries the student's dequant noise
        student_x = teacher_x = None
        if self.perceptual_weight > 0:
            mean, logs = self.student.prior(x, y_onehot=cond)
            latent = gaussian_sample(mean, logs, 1)
            student_x = self.student(z=latent, temperature=0.7, reverse=True, y_onehot=cond)[-1]
            with torch.no_grad():
                teacher_x = self.teacher(z=latent, temperature=0.7, reverse=True, y_onehot=cond)[-1]
        return {"student_nll": student_nll, "student_z": student_z, "teacher_z": teacher_z,
                "student_x": student_x, "teacher_x": teacher_x, "weights": weights, "student_objective": objective}

    @staticmethod
    def _dequantize(x):
        """(x with the dequantisation noise added in place, constant log-det, first SqueezeLayer's output or None):
        one fused kernel when the batch layout allows it (also takes raw uint8 pixels), else the reference's
        uniform_binning_correction followed by the model's own squeeze."""
        from .models.utils import can_fuse_dequant_squeeze, dequantize_and_squeeze, uniform_binning_correction
        if can_fuse_dequant_squeeze(x):
            return dequantize_and_squeeze(x)
        x, ld = uniform_binning_correction(x)
        return x, ld, None

    def _forward_two_streams(self, x, cond, defer=False):
        """Same arithmetic and in-place side effects as the sequential code above (student noise, then teacher noise
        on top, both added to the caller's batch), but the frozen teacher runs on a second stream: its many small
        level-2/3 kernels fill the SMs the student leaves idle. Captured CUDA graphs keep the fork/join. The
        dequantisation kernel hands each model its own squeezed copy of the noised batch, so the teacher's second,
        in-place noise draw cannot disturb the student (no clone of the batch)."""
        main = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream()
        x, ld_s, sq_s = self._dequantize(x)                # student's dequantisation noise, in place
        xs = x if sq_s is not None else x.clone()          # (unfused layout: the student needs its own view of x)
        self._side.wait_stream(main)
        with torch.cuda.stream(self._side), torch.no_grad():
            xt, ld_t, sq_t = self._dequantize(x)           # teacher's noise on top, in place (reference semantics)
            teacher_z, _, _ = self.teacher.flow_from_dequantized(xt, ld_t, cond, _sq0=sq_t)
        if defer:
            student_z, ld, rows, scale = self.student.deferred_objective(xs, ld_s, cond, _sq0=sq_s)
            student_nll = (ld, rows, scale)
        else:
            student_z, student_nll, _ = self.student.flow_from_dequantized(xs, ld_s, cond, _sq0=sq_s)
        main.wait_stream(self._side)
        for i in self.teacher_kd_indices:
            teacher_z[i].record_stream(main)
        return student_z, student_nll, teacher_z

    # ---- pl_module.py:257-320
    def loss(self, out, *args):
        """Multi-level latent MSE, objective, perceptual L1, their weighted sum (times the RICH sample weights) and the
        four batch means: one fused kernel forward, one backward (functional.KdNllLossFn, csrc/loss_optim.cu)."""
        pairs = list(zip(self.student_kd_indices, self.teacher_kd_indices)) if self.kd_weight > 0 else []
        s_levels = [out["student_z"][s] for s, _ in pairs]
        t_levels = [out["teacher_z"][t] for _, t in pairs]
        perc = None
        if self.perceptual_weight > 0:
            d = (out["student_x"] - out["teacher_x"]).abs()
            perc = d.flatten(1).mean(1)
            perc = torch.where(torch.isnan(perc), torch.zeros_like(perc), perc)
        spec = {"teacher": t_levels, "w": (self.nll_weight, self.kd_weight, self.perceptual_weight),
                "sample_w": out["weights"]}
        obj = out.get("student_objective")
        if obj is not None:
            ld, rows, scale = obj
            spec.update(prior=rows, nll_scale=scale)
            means, _, _ = Fn.KdNllLossFn.apply(spec, ld, out["student_z"][-1], perc, *s_levels)
        else:
            means, _, _ = Fn.KdNllLossFn.apply(spec, out["student_nll"], None, perc, *s_levels)
        return {"nll": means[0], "kd": means[1], "perceptual": means[2], "result_loss": means[3]}

    # ---- pl_module.py:322-346
    @torch.no_grad()
    def generate(self, batch):
        name = self.params["data"]["name"]
        if name in TABULAR:
            condition = batch[0]
        elif "drop_weights" not in self.params["data"]:
            _, condition = batch
        else:
            _, condition, _ = batch
        if self.params["student"]["is_1d"] or self.params["student"]["y_condition"]:
            # in 1-D the batch is passed as y_onehot only to size the prior (flows.py:372-376)
            return self.student(reverse=True, y_onehot=condition, temperature=1)[-1]
        return self.student(reverse=True, temperature=1)[-1]

    # ---- pl_module.py:348-363
    def configure_optimizers(self):
        kw = dict(lr=self.params["learning_rate"], weight_decay=self.params["weight_decay"])
        if self.params["optimizer"] == "adam":
            return torch.optim.Adam(self.student.parameters(), **kw)
        if self.params["optimizer"] == "adamax":
            return torch.optim.Adamax(self.student.parameters(), **kw)
        raise NameError("Unknown optimizer name")

    # ---- software pipelining of the frozen teacher (used by train.KDTrainer(pipelined=True))
    def can_stage(self) -> bool:
        """True when the teacher's work for a batch can be done ahead of the student's: 2-D Glow pair with a KD term,
        no perceptual term, unconditioned, fused dequantisation available, fixed-row student prior."""
        return (self.kd_weight > 0 and self.perceptual_weight == 0 and not self.params["student"]["is_1d"]
                and not self.params["teacher"]["is_1d"] and not self.params["student"]["y_condition"]
                and hasattr(self.student, "deferred_objective") and self.student._prior_rows() is not None)

    @torch.no_grad()
    def stage_batch(self, x):
        """Everything of forward() that does not depend on the student's weights, for ONE batch: the student's
        dequantisation noise (in place) and its squeezed copy, then the teacher's noise on top and the frozen teacher's
        forward. Returns (sq_s, ld_s, [teacher taps in KD order]). The teacher is frozen, so running this for batch t+1
        while the student trains on batch t changes no result (pl_module.py:215-227: same draws, same order per batch)."""
        x, ld_s, sq_s = self._dequantize(x)
        assert sq_s is not None, "staging needs the fused dequantisation layout (fp32 / uint8 NCHW, W % 4 == 0)"
        xt, ld_t, sq_t = self._dequantize(x)
        teacher_z, _, _ = self.teacher.flow_from_dequantized(xt, ld_t, None, _sq0=sq_t)
        return sq_s, ld_s, [teacher_z[i] for i in self.teacher_kd_indices]

    def train_on_staged(self, sq_s, ld_s, taps, x_shape):
        """The student's half of training_step for a staged batch: forward from the squeezed noised input, the fused
        loss against the staged teacher taps; returns the same dict as training_step."""
        x_like = torch.empty(x_shape, device="meta")           # placeholder: only its shape is consulted below
        student_z, ld, rows, scale = self.student.deferred_objective(x_like, ld_s, None, _sq0=sq_s)
        t_all = {t: tap for t, tap in zip(self.teacher_kd_indices, taps)}
        out = {"student_nll": None, "student_z": student_z, "teacher_z": t_all, "student_x": None, "teacher_x": None,
               "weights": None, "student_objective": (ld, rows, scale)}
        losses = self.loss(out)
        return {"nll": losses["nll"], "kd": losses["kd"], "perceptual": losses["perceptual"],
                "loss": losses["result_loss"]}

    # ---- pl_module.py:365-382
    def training_step(self, batch, batch_idx=0):
        losses = self.loss(self.forward(batch, _defer_objective=True))
        self.log("train_batch_nll", losses["nll"], on_step=True)
        self.log("train_batch_kd", losses["kd"], on_step=True)
        self.log("train_batch_perceptual", losses["perceptual"], on_step=True)
        self.log("train_batch_loss", losses["result_loss"], on_step=True, prog_bar=True)
        return {"nll": losses["nll"], "kd": losses["kd"], "perceptual": losses["perceptual"],
                "loss": losses["result_loss"]}
