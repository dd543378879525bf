"""KD training driver: what Lightning's loop does around NFModel.training_step in the reference
(train.py:41-58: seed, gradient_clip_val=30, optimiser step), one process per GPU.

The whole step — student forward, teacher forward, multi-level KD loss, backward, gradient clipping and the
optimiser update — is captured once into CUDA graphs and replayed; batches are sharded across ranks and the only
collective is one NCCL all-reduce (average) of the flat student-gradient buffer between the two graphs.
"""
from __future__ import annotations

import os
import typing as tp

import torch
import torch.distributed as dist

from .pl_module import NFModel

GRAD_CLIP = 30.0  # train.py:46 gradient_clip_val


def glow_cfg(image_shape, K, L, hidden, is_1d=False, y_classes=10):
    return dict(image_shape=list(image_shape), hidden_channels=hidden, K=K, L=L, actnorm_scale=1.0,
                flow_permutation="invconv", flow_coupling="affine", LU_decomposed=True, y_classes=y_classes,
                learn_top=False, y_condition=False, is_1d=is_1d)


def kd_config(student, teacher, data="cifar", nll=0.9, kd=0.1, perceptual=0.0, optimizer="adam", lr=5e-4):
    """Same nesting as the reference's Hydra config (conf/config.yaml + conf/training/*.yaml)."""
    return {"data": {"name": data}, "student": student, "teacher": teacher,
            "loss": {"nll": {"weight": nll}, "kd": {"weight": kd, "name": "mse"},
                     "perceptual": {"weight": perceptual, "name": "l1"}},
            "optimizer": optimizer, "learning_rate": lr, "weight_decay": 0.0, "seed": 42}


def allreduce_mean_(grads: tp.Sequence[torch.Tensor], flat: torch.Tensor) -> None:
    """Average `grads` over all ranks IN PLACE with ONE collective: pack into `flat` (one multi-tensor copy), all-reduce,
    unpack. The only communication of the KD step; NCCL over NVLink on GPUs, gloo in the CPU tests."""
    views, off = [], 0
    for g in grads:
        views.append(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
    torch._foreach_copy_(views, list(grads))
    if dist.get_backend() == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG)
    else:   # gloo has no AVG
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(dist.get_world_size())
    torch._foreach_copy_(list(grads), views)


def shard_batch(global_batch: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Equal contiguous shards: per-rank means then average to the global mean (pl_module.py:315-320)."""
    assert global_batch.shape[0] % world == 0, "global batch must divide evenly across ranks"
    n = global_batch.shape[0] // world
    return global_batch[rank * n:(rank + 1) * n]


def randomise_zero_params(model: torch.nn.Module, seed: int, std: float = 0.05):
    """Reference init leaves Conv2dZeros / ActNorm tensors at zero, which turns every coupling into a constant
    (SURVEY §0.5); benchmarks and tests re-draw them N(0, std^2) so all kernels do real work."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in model.parameters():
            if p.abs().max() == 0:
                p.copy_((torch.randn(p.shape, generator=g) * std).to(p.device))


class FlatAdam:
    """torch.nn.utils.clip_grad_norm_(max_norm) + torch.optim.Adam / Adamax (pl_module.py:348-363, train.py:46) on FLAT
    fp32 buffers, two launches per step (csrc/loss_optim.cu): every parameter's storage is re-pointed into one
    contiguous buffer `p` (16-byte aligned slices), the gradients are gathered into `g` (the buffer the all-reduce
    runs on), and the moments live in `m` / `v`. The step counter is a device int so a captured graph advances it."""

    ALIGN = 64   # floats: every slice starts on a 256-byte boundary

    def __init__(self, params: tp.Sequence[torch.nn.Parameter], lr: float, weight_decay: float = 0.0,
                 kind: str = "adam", max_norm: float = GRAD_CLIP, betas=(0.9, 0.999), eps: float = 1e-8):
        from . import ops
        self.ops = ops
        self.params = list(params)
        dev = self.params[0].device
        self.offsets, off = [], 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.n = off
        self.p = torch.zeros(off, device=dev)
        self.g = torch.zeros(off, device=dev)
        self.m = torch.zeros(off, device=dev)
        self.v = torch.zeros(off, device=dev)
        self.step_count = torch.zeros(1, device=dev, dtype=torch.int32)
        self.partials = torch.zeros(ops.optim_partials(), device=dev)
        self.grad_norm = torch.zeros(1, device=dev)
        self.lr, self.wd, self.max_norm, self.betas, self.eps = lr, weight_decay, max_norm, betas, eps
        if kind not in ("adam", "adamax"):
            raise NameError("Unknown optimizer name")
        self.adamax = kind == "adamax"
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                view = self.p[o:o + p.numel()].view_as(p)
                view.copy_(p)
                p.data = view                    # the module's kernels now read the flat buffer directly
        self.g_views = [self.g[o:o + p.numel()].view_as(p) for p, o in zip(self.params, self.offsets)]

    def gather_grads(self):
        """p.grad (fresh tensors produced by backward) -> slices of the flat gradient buffer, one multi-tensor copy."""
        torch._foreach_copy_(self.g_views, [p.grad for p in self.params])

    def step(self):
        o = self.ops
        o.grad_sqnorm(self.g, self.partials, self.step_count)
        o.adam_step(self.p, self.g, self.m, self.v, self.partials, self.step_count, self.max_norm, self.lr,
                    self.betas[0], self.betas[1], self.eps, self.wd, self.adamax, self.grad_norm)

    def reset(self):
        self.m.zero_(); self.v.zero_(); self.g.zero_(); self.step_count.zero_()


class KDTrainer:
    """Owns the NFModel, its optimiser, the static device buffers and the captured graphs."""

    def __init__(self, config: dict, batch_shape: tp.Sequence[int], device: torch.device, use_graphs: bool = True,
                 seed: int = 42, input_dtype: torch.dtype = torch.float32, pipelined: bool = False):
        """input_dtype=torch.uint8: image batches arrive as raw pixels (a quarter of the host-to-device bytes); the
        first kernel of the step does the reference's preprocess together with the noise and the first squeeze.

        pipelined=True (2-D Glow pairs, see NFModel.can_stage): the FROZEN teacher's forward for the batch in `self.x`
        runs concurrently with the student's forward / backward / update on the PREVIOUS batch (teacher prefetch of
        depth one). A step still consumes one batch and makes one optimiser update; the updates are the same as without
        pipelining (the teacher does not depend on the student), `self.losses` after `step()` are those of the batch
        handed to the previous `step()`. The dependent launch chain of a step is then max(teacher, student) instead of
        teacher followed by the student's backward, which is what bounds the step at small per-GPU batches."""
        self.device = device
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        torch.manual_seed(seed)                       # identical initial weights on every rank (train.py:41)
        self.module = NFModel(config)
        randomise_zero_params(self.module.student, seed + 1)
        if self.module.teacher is not None:
            randomise_zero_params(self.module.teacher, seed + 2)
        self.module.to(device)
        self.is_1d = config["student"]["is_1d"]
        params = [p for p in self.module.student.parameters() if p.requires_grad]
        self.params = params
        # flat parameter / gradient / moment buffers: the all-reduce, the clipping norm and the optimiser all run on
        # them; the gradients themselves are produced fresh by backward (p.grad = None before it), so autograd steals
        # them instead of launching one "+=" kernel per parameter
        self.opt = FlatAdam(params, lr=config["learning_rate"], weight_decay=config["weight_decay"],
                            kind=config["optimizer"])
        self.x = torch.zeros(*batch_shape, device=device, dtype=input_dtype)          # static input buffer
        self.losses = torch.zeros(4, device=device)                 # nll, kd, perceptual, loss
        self.use_graphs = use_graphs
        self.pipelined = bool(pipelined) and self.module.can_stage()
        self.staged = None          # (sq_s, ld_s, taps) of the batch the student trains on next
        self._side = torch.cuda.Stream(device=device) if self.pipelined else None
        self.g_fb: tp.Optional[torch.cuda.CUDAGraph] = None
        self.g_opt: tp.Optional[torch.cuda.CUDAGraph] = None
        torch.manual_seed(seed + 1000 + (dist.get_rank() if dist.is_initialized() else 0))  # per-rank noise stream

    # ---- the two halves of a step
    def _forward_backward(self):
        for p in self.params:
            p.grad = None
        if self.pipelined:
            return self._forward_backward_pipelined()
        batch = [self.x] if self.is_1d else [self.x, None]
        out = self.module.training_step(batch, 0)
        out["loss"].backward()
        self.losses.copy_(torch.stack([out["nll"], out["kd"], out["perceptual"], out["loss"]]).detach())
        self.opt.gather_grads()

    def _stage_next(self):
        """Teacher side of the batch in self.x as fresh tensors (the static `staged` buffers are allocated on first use)."""
        sq, ld, taps = self.module.stage_batch(self.x)
        if self.staged is None:
            self.staged = (torch.empty_like(sq), torch.empty_like(ld), [torch.empty_like(t) for t in taps])
        return sq, ld, taps

    def _forward_backward_pipelined(self):
        """student(batch t) on the main stream  ||  teacher(batch t+1, in self.x) on the side stream; then t+1's staged
        tensors replace t's. The first call only stages (prime())."""
        main = torch.cuda.current_stream()
        self._side.wait_stream(main)
        with torch.cuda.stream(self._side):
            nxt = self._stage_next()
        sq, ld, taps = self.staged
        out = self.module.train_on_staged(sq, ld, taps, tuple(self.x.shape))
        out["loss"].backward()
        self.losses.copy_(torch.stack([out["nll"], out["kd"], out["perceptual"], out["loss"]]).detach())
        self.opt.gather_grads()
        main.wait_stream(self._side)
        for t in (nxt[0], nxt[1], *nxt[2]):
            t.record_stream(main)
        sq.copy_(nxt[0]); ld.copy_(nxt[1])
        torch._foreach_copy_(taps, nxt[2])

    def prime(self):
        """Pipelined mode: stage the batch currently in self.x (no training). Call once before the first step."""
        if not self.pipelined:
            return
        sq, ld, taps = self._stage_next()
        self.staged[0].copy_(sq); self.staged[1].copy_(ld)
        torch._foreach_copy_(self.staged[2], taps)

    def _clip_and_update(self):
        self.opt.step()

    def _allreduce(self):
        """The only collective of the KD step: average the flat student gradient over the ranks."""
        if self.world > 1:
            dist.all_reduce(self.opt.g, op=dist.ReduceOp.AVG)

    def flat_grad_view(self) -> torch.Tensor:
        """The flat gradient buffer of the last step (after the all-reduce, before clipping)."""
        return self.opt.g

    def reset_state(self, student_state: tp.Optional[dict] = None):
        """Back to a given student state_dict with fresh optimiser state (warm-up / capture steps must not count)."""
        from . import functional as Fn
        if student_state is not None:
            self.module.student.load_state_dict(student_state)     # in-place copies: the flat views stay
        self.opt.reset()
        Fn.bump_param_epoch()

    def warmup(self, iters: int = 3):
        """Eager steps on a side stream (builds kernels' attributes, optimiser state), then graph capture."""
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self.prime()
            for _ in range(iters):
                self._forward_backward()
                self._allreduce()
                self._clip_and_update()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        if not self.use_graphs:
            return
        # One rank: the whole step (forward, backward, clip + optimiser) is ONE graph. Several ranks: two graphs with
        # the NCCL all-reduce launched eagerly in between (the configuration measured at 2 and 8 GPUs).
        # NFK_GRAPH_NCCL=1 captures the collective inside a single graph instead — EXPERIMENTAL and off by default: the
        # one 2-GPU attempt made with it in round 2 did not finish inside its time limit and was not investigated.
        one_graph = self.world == 1 or os.environ.get("NFK_GRAPH_NCCL", "0") == "1"
        if one_graph:
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._forward_backward()
                    self._allreduce()
                    self._clip_and_update()
                self.g_fb, self.g_opt = g, None
            except Exception as e:   # noqa: BLE001  (capture of the collective not supported by this stack)
                if self.world == 1:
                    raise
                import warnings
                warnings.warn(f"single-graph capture with the NCCL all-reduce failed ({e}); using two graphs")
                torch.cuda.synchronize()
                one_graph = False
        if not one_graph:
            self.g_fb = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g_fb):
                self._forward_backward()
            self.g_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g_opt, pool=self.g_fb.pool()):
                self._clip_and_update()
        self.one_graph = one_graph
        torch.cuda.synchronize()

    def step_device(self):
        """One training step on whatever self.x currently holds (inputs already resident in HBM)."""
        from . import functional as Fn
        if self.g_fb is not None:
            self.g_fb.replay()
            if self.g_opt is not None:
                self._allreduce()
                self.g_opt.replay()
        else:
            self._forward_backward()
            self._allreduce()
            self._clip_and_update()
        Fn.bump_param_epoch()     # parameter values changed behind autograd's back: drop cached derived operands

    def step(self, host_batch: torch.Tensor) -> torch.Tensor:
        """Public end-to-end step: pinned host batch -> device, train, loss scalars back on the host."""
        self.x.copy_(host_batch, non_blocking=True)
        self.step_device()
        return self.losses.cpu()


def init_distributed() -> tp.Tuple[int, int, torch.device]:
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, torch.device("cuda", local)
