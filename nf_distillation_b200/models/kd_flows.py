"""Drop-in mirror of the reference's models/kd_flows.py: the KD variants return EVERY layer's output so the
training module can pick its distillation taps (reference /root/reference/models/kd_flows.py)."""
from __future__ import annotations

import logging
import math
import typing as tp

import torch

from .. import functional as Fn
from .flows import FlowNet, FlowStep, Glow
from .layers import Split2d, _require_cuda
from .utils import can_fuse_dequant_squeeze, dequantize_and_squeeze, uniform_binning_correction

logger = logging.getLogger(__name__)


class FlowNetGetAllOutputs(FlowNet):
    """encode/decode return the list of all layer outputs (kd_flows.py:15-73)."""

    def encode(self, z, y_onehot=None, logdet=0.0, _sq0=None):
        all_outputs = []
        with Fn.use_prep(Fn.prepare_steps(self.layers, False)):   # one K0 launch for every trainable step
            for z, logdet in self._encode_layers(z, y_onehot, logdet, _sq0):
                all_outputs.append(z)
        return all_outputs, logdet

    def decode(self, z, y_onehot=None, temperature=None):
        all_outputs = []
        with Fn.use_prep(Fn.prepare_steps(self.layers, True)):
            for layer in reversed(self.layers):
                if isinstance(layer, Split2d):
                    z, _ = layer(z, logdet=0, reverse=True, temperature=temperature)
                else:
                    z, _ = layer(z, y_onehot=y_onehot, logdet=0, reverse=True)
                all_outputs.append(z)
        return all_outputs


class GlowGetAllOutputs(Glow):
    """Glow whose forward returns (list of all layer outputs, bpd [B], y_logits) (kd_flows.py:76-152).

    Like the reference, the constructor first builds a plain FlowNet (inside Glow.__init__) and then replaces it,
    so the global RNG is consumed twice and seed-for-seed initial weights match the reference."""

    def __init__(self, image_shape, hidden_channels, K, L, actnorm_scale, flow_permutation, flow_coupling,
                 LU_decomposed, y_classes, learn_top, y_condition, is_1d=False):
        super().__init__(image_shape, hidden_channels, K, L, actnorm_scale, flow_permutation, flow_coupling,
                         LU_decomposed, y_classes, learn_top, y_condition, is_1d=is_1d)
        self.flow = FlowNetGetAllOutputs(image_shape=image_shape, hidden_channels=hidden_channels, K=K, L=L,
                                         actnorm_scale=actnorm_scale, flow_permutation=flow_permutation,
                                         flow_coupling=flow_coupling, LU_decomposed=LU_decomposed, is_1d=is_1d,
                                         condition_features=y_classes if y_condition else 0)

    def normal_flow(self, x, y_onehot):
        _require_cuda(x, "GlowGetAllOutputs")
        sq0 = None
        if self.is_1d:
            logdet = torch.zeros(x.shape[0], device=x.device, dtype=torch.float32)
        elif can_fuse_dequant_squeeze(x):
            x, logdet, sq0 = dequantize_and_squeeze(x)     # noise + first squeeze (+ uint8 preprocess): one kernel
        else:
            x, logdet = uniform_binning_correction(x)
        return self.flow_from_dequantized(x, logdet, y_onehot, _sq0=sq0)

    def flow_from_dequantized(self, x, logdet, y_onehot=None, _sq0=None):
        """Everything of normal_flow after the (in-place) dequantisation; NFModel uses it to run the teacher and the
        student concurrently on two streams while keeping the reference's in-place noise semantics. `_sq0`: the first
        SqueezeLayer's output when the dequantisation kernel already produced it (x is then only read for its shape)."""
        z, logdet = self.flow(x, y_onehot=y_onehot, logdet=logdet, reverse=False, _sq0=_sq0)
        last_z = z[-1]
        bpd = self._objective(x, last_z, logdet, y_onehot)
        if self.y_condition:
            pooled = last_z if self.is_1d else last_z.mean(dim=[2, 3])
            y_logits = self.project_class(pooled)
        else:
            y_logits = None
        return z, bpd, y_logits

    def deferred_objective(self, x, logdet, y_onehot=None, _sq0=None):
        """flow_from_dequantized WITHOUT the prior / bits-per-dim reduction: returns (z list, logdet [B],
        (mean_row, logs_row), nll_scale) so that NFModel.loss can fold the objective (kd_flows.py:134-150) into its
        single fused loss kernel (functional.KdNllLossFn); None when the prior is not a batch-independent row
        (learn_top / y_condition), in which case the caller uses the ordinary forward."""
        rows = self._prior_rows()
        if rows is None:
            return None
        z, logdet = self.flow(x, y_onehot=y_onehot, logdet=logdet, reverse=False, _sq0=_sq0)
        scale = 1.0 if self.is_1d else 1.0 / (math.log(2.0) * x.shape[1] * x.shape[2] * x.shape[3])
        return z, logdet, rows, scale


def create_glow_model(config: tp.Dict[str, tp.Any]) -> GlowGetAllOutputs:
    """kd_flows.py:155-159: build the model and mark every ActNorm as initialised."""
    model = GlowGetAllOutputs(**config)
    model.set_actnorm_init()
    return model


def inherit_permutation_matrix(student, teacher, student_kd_indices, teacher_kd_indices):
    """Compose teacher permutation matrices between KD taps into the student's (kd_flows.py:162-179)."""
    k = 0
    acc = None
    for t_id, t_layer in enumerate(teacher.flow.layers):
        if t_id == teacher_kd_indices[k]:
            student.flow.layers[student_kd_indices[k]].invconv.p = acc @ t_layer.invconv.p
            k += 1
            acc = None
        elif isinstance(t_layer, FlowStep):
            acc = t_layer.invconv.p if acc is None else acc @ t_layer.invconv.p
