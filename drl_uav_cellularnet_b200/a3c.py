"""The actor-critic learner on top of the batched env step path -- host-side mirror of the
reference's ``ACNet`` / ``Worker`` (main.py:43-271) for ``netType='MLP'``.

What changes against the reference
  * The first layer never sees the dense 50 000-wide observation: the env kernel emits the ~44 non-zero cells as flat
    indices (``obs_idx``) and ``uavnet_sparse_fwd`` / ``uavnet_sparse_bwd`` (include/uavnet.h) do the gather-sum and
    the scatter-add.  Actor and critic first layers are stored side by side in one ``[N_S, 2*200]`` matrix so that one
    gather feeds both.  The 200x200 / 200x625 / 200x1 layers and all their gradients run on the tcgen05 tensor cores
    through ``uavnet_gemm`` (dense.py, csrc/tc_gemm.cuh) with bias / relu6 / relu6' / value-head / bias-gradient
    epilogues fused in; ``precision='fp32'`` is 3xTF32 (fp32-class accuracy), ``'tf32'`` one MMA per k-step.
  * Four asynchronous worker threads pushing gradients into a shared net (Hogwild, main.py:85-86,159-163) become one
    synchronous batch: every rank rolls its E envs UPDATE_GLOBAL_ITER steps, the loss is the mean over all samples
    (like the reference's own synchronous variant, a2c_single_thread.py:153-186), gradients are summed over ranks with
    one NCCL all-reduce of the flat buffer and both RMSProp optimisers (same hyper-parameters) run as one fused pass.
  * Backward is written out by hand (formulas of main.py:64-78); tests/test_gpu_a3c.py checks it against
    torch.autograd on the dense restatement of the same graph.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np
import torch

from . import _native as N
from . import dense

# the reference's hyper-parameters (main.py:19-27)
UPDATE_GLOBAL_ITER = 10
GAMMA = 0.9
ENTROPY_BETA = 0.001
LR_A = 0.0001
LR_C = 0.0001
TENSOR_SEED = 6
HIDDEN = 200
# tf.train.RMSPropOptimizer defaults (TF 1.x): decay, momentum (0), epsilon; the `rms` slot starts at ones
RMS_DECAY, RMS_EPS = 0.9, 1e-10


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _relu6(x):
    return torch.clamp(x, 0.0, 6.0)


class _ForeignCudaBuffer:
    """float32 device memory owned by libuavenv (uavnet_p2p_alloc), exposed to torch through __cuda_array_interface__"""

    def __init__(self, ptr: int, n: int):
        self.ptr, self.n = ptr, n
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2, "strides": None}


class ACNet:
    """Actor 50000->200->200->N_A softmax and critic 50000->200->200->1 (main.py:143-156) on sparse observations.

    All parameters live in ONE flat float32 buffer (``self.flat``; gradient ``self.grad`` and RMSProp slot ``self.ms``
    have the same layout), segments 16-byte aligned:
        W1 [N_S, 2H] (cols :H = actor 'la', H: = critic 'lc') | b1 [2H] | Wa2 [H,H] | ba2 [H] | Wa3 [H,N_A] | ba3 [N_A]
        | Wc2 [H,H] | bc2 [H] | Wc3 [H,1] | bc3 [1]
    """

    def __init__(self, n_s: int, n_a: int, device, hidden: int = HIDDEN, seed: int = TENSOR_SEED, precision: str = "fp32"):
        self.n_s, self.n_a, self.h = int(n_s), int(n_a), int(hidden)
        if precision not in dense.PRECISIONS:
            raise ValueError("precision must be one of %s" % sorted(dense.PRECISIONS))
        self.precision = precision
        # first-layer weight gradient: "gather" (counting sort + per-row sums) or "scatter" (float REDs); measured in
        # profiles/r2/NOTES.md section 5
        self.overlap_chains = os.environ.get("UAVNET_OVERLAP_CHAINS", "1") != "0"
        self.sparse_bwd = os.environ.get("UAVNET_SPARSE_BWD", "gather")
        self.sparse_bwd_passes = int(os.environ.get("UAVNET_BWD_PASSES", "2"))
        self.ld_a = (self.n_a + 3) // 4 * 4          # rows of Wa3 / dz padded to 16 bytes (625 -> 628): vector staging
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("ACNet needs a CUDA device (sm_100a); there is no CPU fallback")
        self._lib = N.lib()
        H = self.h
        shapes = [("W1", (n_s, 2 * H)), ("b1", (2 * H,)), ("Wa2", (H, H)), ("ba2", (H,)), ("Wa3", (H, n_a)), ("ba3", (n_a,)),
                  ("Wc2", (H, H)), ("bc2", (H,)), ("Wc3", (H, 1)), ("bc3", (1,))]
        self.segments, off = {}, 0
        for name, shp in shapes:
            stored = (shp[0], self.ld_a) if name == "Wa3" else shp          # padding columns stay zero for ever
            n = int(np.prod(stored))
            self.segments[name] = (off, n, stored, shp)
            off += (n + 3) // 4 * 4
        self.n_flat = off
        self.flat = torch.zeros(off, dtype=torch.float32, device=self.device)
        self.grad = torch.zeros_like(self.flat)
        self.ms = torch.ones_like(self.flat)                # TF1 RMSProp: rms slot initialised to ones
        self._bind_views()
        # tf.random_normal_initializer(0., .1, seed) kernels, zero biases (main.py:145-153).  The TF stream itself is
        # not reproducible here (TensorFlow absent); the distribution is.
        gen = torch.Generator(device="cpu").manual_seed(int(seed))
        for k in ("W1", "Wa2", "Wa3", "Wc2", "Wc3"):
            self.p[k].copy_(torch.randn(self.segments[k][3], generator=gen) * 0.1)
        self.n_params = sum(int(np.prod(shp)) for _, _, _, shp in self.segments.values())
        self._scratch = {}
        # [N, K] copies of the forward weights: both operands of the forward GEMMs are then row-major and go through TMA
        self.pt = {"Wa2": torch.empty((H, H), dtype=torch.float32, device=self.device),
                   "Wa3": torch.empty((self.n_a, H), dtype=torch.float32, device=self.device),
                   "Wc2": torch.empty((H, H), dtype=torch.float32, device=self.device)}
        self._pt_seen = None

    def _sync_transposed(self):
        """Refresh the transposed weight copies if the parameters changed (torch's version counter sees in-place edits
        through any view; ``apply_grads`` / ``enable_p2p`` mark the raw-pointer writes)."""
        seen = (self.flat.data_ptr(), self.flat._version)
        if self._pt_seen != seen:
            for k, t in self.pt.items():
                t.copy_(self.p[k].t())
            self._pt_seen = (self.flat.data_ptr(), self.flat._version)

    def _bind_views(self):
        def views(buf):
            out = {}
            for k, (o, n, stored, shp) in self.segments.items():
                v = buf[o:o + n].view(stored)
                out[k] = v[:, :shp[1]] if stored != shp else v
            return out
        self.p, self.g = views(self.flat), views(self.grad)

    def _buf(self, name: str, shape, dtype=torch.float32) -> torch.Tensor:
        """persistent scratch (stable pointers: CUDA-graph friendly, no allocator traffic per step)"""
        key = (name, tuple(shape), dtype)
        t = self._scratch.get(key)
        if t is None:
            t = self._scratch[key] = torch.empty(shape, dtype=dtype, device=self.device)
        return t

    def _side_stream(self, i: int = 0):
        if getattr(self, "_sides", None) is None:
            self._sides = {}
        if i not in self._sides:
            self._sides[i] = torch.cuda.Stream(device=self.device)
        return self._sides[i]

    def _gemm(self, A, B, out=None, **kw):
        return dense.gemm(A, B, out, precision=self.precision, **kw)

    # ---- forward ------------------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def first_layer(self, idx: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """relu6(s @ [la | lc] + b) for count vectors given as flat indices idx int32 [M, K] -> [M, 2H]."""
        assert idx.dtype == torch.int32 and idx.is_cuda and idx.is_contiguous() and idx.dim() == 2
        M, K = idx.shape
        if out is None:
            out = torch.empty((M, 2 * self.h), dtype=torch.float32, device=self.device)
        rc = self._lib.uavnet_sparse_fwd(_ptr(idx), M, K, self.n_s, _ptr(self.p["W1"]), _ptr(self.p["b1"]), 2 * self.h,
                                         _ptr(out), 1, self._stream())
        if rc:
            raise RuntimeError("uavnet_sparse_fwd failed (%d)" % rc)
        return out

    def forward(self, idx: torch.Tensor, want: str = "both", out: Optional[dict] = None):
        """-> (a_prob [M, N_A] or None, v [M] or None, cache).  `out`: preallocated h1 / h2a / prob buffers (the
        trainer's rollout storage) to write the activations into."""
        H, p = self.h, self.p
        self._sync_transposed()
        h1 = self.first_layer(idx, None if out is None else out["h1"])
        cache = {"idx": idx, "h1": h1}
        prob = v = None
        if want in ("both", "actor"):
            h2a = self.actor_hidden(h1, None if out is None else out["h2a"])
            # softmax through uavnet_softmax_sample (probabilities only): the same kernel the rollout uses
            prob, _ = self.sample_head(h2a, 0, 0, None, 0, prob_out=None if out is None else out["prob"], want_action=False)
            cache.update(h2a=h2a, prob=prob)
        if want in ("both", "critic"):
            h2c, v = self.critic_head(h1)
            cache.update(h2c=h2c, v=v)
        return prob, v, cache

    def actor_hidden(self, h1: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """h2a = relu6(h1[:, :H] @ Wa2 + ba2) (main.py:148)"""
        self._sync_transposed()
        return self._gemm(h1[:, :self.h], self.pt["Wa2"], out, b_trans=True, bias=self.p["ba2"], relu6=True)

    def critic_head(self, h1: torch.Tensor, h2c_out: Optional[torch.Tensor] = None, v_out: Optional[torch.Tensor] = None):
        """h2c = relu6(h1[:, H:] @ Wc2 + bc2), v = h2c @ Wc3 + bc3 (main.py:152-153) in one launch: the value head is a
        row dot in the epilogue of the second layer -> (h2c [M, H], v [M])"""
        H, p = self.h, self.p
        M = h1.shape[0]
        v = v_out if v_out is not None else torch.empty(M, dtype=torch.float32, device=self.device)
        self._sync_transposed()
        h2c = self._gemm(h1[:, H:], self.pt["Wc2"], h2c_out, b_trans=True, bias=p["bc2"], relu6=True, dot_w=p["Wc3"].view(-1),
                         dot_b=p["bc3"], dot_out=v)
        return h2c, v

    def sample_head(self, h2a: torch.Tensor, seed: int, row_offset: int, counter_dev: Optional[torch.Tensor], counter_add: int,
                    prob_out: Optional[torch.Tensor] = None, action_out: Optional[torch.Tensor] = None,
                    logits_out: Optional[torch.Tensor] = None, want_action: bool = True):
        """logits = h2a @ Wa3 + ba3, then softmax + np.random.choice(p=a_prob) in one kernel (uavnet_softmax_sample,
        Philox keyed by (seed, row_offset + row, *counter_dev + counter_add)) -> (prob [M, N_A], action int64 [M])"""
        M = h2a.shape[0]
        self._sync_transposed()
        logits = self._gemm(h2a, self.pt["Wa3"], logits_out if logits_out is not None else self._buf("logits", (M, self.n_a)),
                            b_trans=True, bias=self.p["ba3"])
        prob = prob_out if prob_out is not None else torch.empty_like(logits)
        action = None
        if want_action:
            action = action_out if action_out is not None else torch.empty(M, dtype=torch.int64, device=self.device)
        rc = self._lib.uavnet_softmax_sample(_ptr(logits), M, self.n_a, int(seed), int(row_offset), _ptr(counter_dev),
                                             int(counter_add), _ptr(prob), _ptr(action), self._stream())
        if rc:
            raise RuntimeError("uavnet_softmax_sample failed (%d)" % rc)
        return prob, action

    def choose_action(self, idx: torch.Tensor, seed: int = 0, row_offset: int = 0) -> torch.Tensor:
        """np.random.choice(N_A, p=a_prob) per env (main.py:165-169) -> int64 [M]: softmax + inverse-CDF draw in one kernel
        (uavnet_softmax_sample), one Philox uniform per row keyed by (seed, row_offset + row, number of this call)"""
        self._sync_transposed()
        h2a = self.actor_hidden(self.first_layer(idx))
        self._choose_calls = getattr(self, "_choose_calls", 0) + 1
        return self.sample_head(h2a, seed, row_offset, None, self._choose_calls)[1]

    def greedy_action(self, idx: torch.Tensor) -> torch.Tensor:
        """tf.argmax(a_prob, 1) of the evaluation driver (main_test.py:68)"""
        prob, _, _ = self.forward(idx, "actor")
        return prob.argmax(dim=1)

    def value(self, idx: torch.Tensor) -> torch.Tensor:
        return self.forward(idx, "critic")[1]

    # ---- losses + gradients (main.py:64-78), accumulated into self.grad -------------------------------------
    def accumulate_grads(self, idx: torch.Tensor, a_his: torch.Tensor, v_target: torch.Tensor, saved: Optional[dict] = None,
                         early_push=None):
        """Adds d(a_loss)/d(actor params) and d(c_loss)/d(critic params) for the batch to ``self.grad``;
        returns (a_loss, c_loss) as 0-d tensors.  a_loss = mean(-(log(pi(a)+1e-5) * sg(td) + beta * H)),
        c_loss = mean(td^2), td = v_target - v.  `saved`: activations kept from the rollout (h1, h2a, prob -- the
        parameters do not change between the rollout and its update), else they are recomputed from idx.  early_push: called
        once the actor half of the first layer's gradient is complete (split peer-memory push)."""
        H, p, g = self.h, self.p, self.g
        M = idx.shape[0]
        if saved is not None:
            h1, h2a, prob = saved["h1"], saved["h2a"], saved["prob"]
        else:
            prob, _, c = self.forward(idx, "actor")
            h1, h2a = c["h1"], c["h2a"]
        gemm = self._gemm
        side = ws = None
        if self.sparse_bwd == "gather":
            # the bucketing of the first layer's (sample, slot) pairs needs the indices only: it runs on a side stream under
            # the dense layers' backward products and is joined right before the sums
            nb = int(self._lib.uavnet_sparse_bwd_gather_workspace(M, idx.shape[1], self.n_s))
            ws = self._buf("bwd_ws", (nb // 4,), torch.int32)
            cur = torch.cuda.current_stream(self.device)
            side = self._side_stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                rc = self._lib.uavnet_sparse_bwd_gather_prepare(_ptr(idx), M, idx.shape[1], self.n_s, _ptr(ws), self._stream())
            if rc:
                raise RuntimeError("uavnet_sparse_bwd_gather_prepare failed (%d)" % rc)
        h2c, v = self.critic_head(h1, self._buf("h2c", (M, H)), self._buf("v", (M,)))
        # td = v_target - v, dv = -2 td / M and c_loss = mean(td^2) in one launch (uavnet_critic_td)
        td, dv1, loss2 = self._buf("td", (M,)), self._buf("dv", (M,)), self._buf("loss2", (2,))
        vt = v_target.contiguous()
        rc = self._lib.uavnet_critic_td(_ptr(vt), _ptr(v), M, _ptr(td), _ptr(dv1), _ptr(loss2), self._stream())
        if rc:
            raise RuntimeError("uavnet_critic_td failed (%d)" % rc)
        dv = dv1.unsqueeze(1)                                              # [M, 1]
        dpre1 = self._buf("dpre1", (M, 2 * H))
        # -- critic: every weight gradient is x^T @ dy accumulated into the flat gradient buffer, its bias gradient the
        # -- column sums of dy from the same pass; every data gradient is (dy @ W^T) * relu6'(layer output).  The critic's
        # -- backward chain and the actor's touch disjoint buffers once td exists: the critic's runs on a second stream --
        main = torch.cuda.current_stream(self.device)
        cstream = self._side_stream(1) if self.overlap_chains else main
        if cstream is not main:
            cstream.wait_stream(main)
        with torch.cuda.stream(cstream):
            gemm(h2c, dv, g["Wc3"], a_trans=True, accumulate=True, colsum=g["bc3"])
            dpre2c = self._buf("dpre2c", (M, H))                           # (dv (x) wc3) * relu6'(h2c)
            rc = self._lib.uavnet_rank1_mask(_ptr(dv), _ptr(p["Wc3"]), _ptr(h2c), M, H, _ptr(dpre2c), self._stream())
            if rc:
                raise RuntimeError("uavnet_rank1_mask failed (%d)" % rc)
            gemm(h1[:, H:], dpre2c, g["Wc2"], a_trans=True, accumulate=True, colsum=g["bc2"])
            gemm(dpre2c, p["Wc2"], dpre1[:, H:], b_trans=True, mask_src=h1[:, H:], out_colsum=g["b1"][H:])
        # -- actor: d(a_loss)/d(logits) in one fused pass over the softmax output --
        dz = self._buf("dz", (M, self.ld_a))[:, :self.n_a]
        loss_row = self._buf("loss_row", (M,))
        rc = self._lib.uavnet_actor_head_bwd(_ptr(prob), _ptr(a_his), _ptr(td), M, self.n_a, ENTROPY_BETA, _ptr(dz), self.ld_a,
                                             _ptr(loss_row), self._stream())
        if rc:
            raise RuntimeError("uavnet_actor_head_bwd failed (%d)" % rc)
        rc = self._lib.uavnet_mean_rows(_ptr(loss_row), M, _ptr(loss2), self._stream())       # a_loss = mean of the rows
        if rc:
            raise RuntimeError("uavnet_mean_rows failed (%d)" % rc)
        gemm(h2a, dz, g["Wa3"], a_trans=True, accumulate=True, colsum=g["ba3"])
        dpre2a = gemm(dz, p["Wa3"], self._buf("dpre2a", (M, H)), b_trans=True, mask_src=h2a)
        gemm(h1[:, :H], dpre2a, g["Wa2"], a_trans=True, accumulate=True, colsum=g["ba2"])
        gemm(dpre2a, p["Wa2"], dpre1[:, :H], b_trans=True, mask_src=h1[:, :H], out_colsum=g["b1"][:H])   # + first-layer bias gradient
        if cstream is not main:
            main.wait_stream(cstream)
        if self.sparse_bwd == "gather":
            # rows bucketed by a counting sort, then one plain sum per row and column half (no float atomics)
            torch.cuda.current_stream(self.device).wait_stream(side)
            if early_push is not None and H % 4 == 0:
                # the actor half first; its push (early_push: a callable, e.g. ACNet.push_early) goes out under the critic half
                rc = self._lib.uavnet_sparse_bwd_gather_apply_cols(M, idx.shape[1], self.n_s, _ptr(dpre1), 2 * H, 0, H, _ptr(g["W1"]),
                                                                   _ptr(ws), self._stream())
                if rc:
                    raise RuntimeError("uavnet_sparse_bwd_gather_apply_cols failed (%d)" % rc)
                early_push()
                rc = self._lib.uavnet_sparse_bwd_gather_apply_cols(M, idx.shape[1], self.n_s, _ptr(dpre1), 2 * H, H, H, _ptr(g["W1"]),
                                                                   _ptr(ws), self._stream())
            else:
                rc = self._lib.uavnet_sparse_bwd_gather_apply(M, idx.shape[1], self.n_s, _ptr(dpre1), 2 * H, _ptr(g["W1"]), _ptr(ws),
                                                              self.sparse_bwd_passes if (2 * H // 4) % self.sparse_bwd_passes == 0 else 1,
                                                              self._stream())
        else:
            rc = self._lib.uavnet_sparse_bwd(_ptr(idx), M, idx.shape[1], self.n_s, _ptr(dpre1), 2 * H, _ptr(g["W1"]),
                                             self._stream())
        if rc:
            raise RuntimeError("uavnet_sparse_bwd failed (%d)" % rc)
        return loss2[0], loss2[1]                                          # (a_loss, c_loss): views of a persistent buffer

    # ---- the "push" (main.py:85-86,159-160): all-reduce + both RMSProp optimisers in one pass ---------------
    def apply_grads(self, lr: float = LR_A, world_size: int = 1):
        """grad /= world_size (after the caller's all-reduce), RMSProp step on every parameter, grad = 0.
        With ``enable_p2p()`` the all-reduce is not needed: one peer-memory kernel does reduce + step + broadcast."""
        self._pt_seen = None                       # the kernels below write the parameters through raw pointers
        if getattr(self, "_p2p", None):
            return self._apply_grads_p2p(lr)
        rc = self._lib.uavnet_rmsprop(_ptr(self.flat), _ptr(self.grad), _ptr(self.ms), self.n_flat, lr, RMS_DECAY, RMS_EPS,
                                      1.0 / world_size, 1, self._stream())
        if rc:
            raise RuntimeError("uavnet_rmsprop failed (%d)" % rc)

    # ---- the push over NVLink peer memory: reduce-scatter + RMSProp + all-gather in ONE kernel per rank ----------
    def enable_p2p(self, split: bool = False):
        """Move the flat parameter / gradient buffers into IPC-shareable allocations, exchange their handles between
        the ranks' processes and map every peer's buffers (cudaIpcOpenMemHandle, peer access over NVLink).  From then
        on ``apply_grads`` is ``uavnet_p2p_push``: one peer-memory kernel per rank does reduce-scatter + RMSProp +
        all-gather, ordered between the ranks by flag words in peer memory -- no NCCL call in the push, nothing for the
        host to wait for, replayable in a CUDA graph.  World size 1 (no process group) takes the same code path with
        the local buffers only."""
        import torch.distributed as dist
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank() if world > 1 else 0
        if world > 8:
            raise ValueError("uavnet_p2p_push supports up to 8 ranks")
        vp = C.c_void_p
        bufs, handles = [], []
        for nbytes in (self.n_flat * 4, self.n_flat * 4, 256):        # 0: gradients, 1: parameters, 2: flag words (zeroed)
            ptr, hdl = vp(), (C.c_uint8 * 64)()
            rc = self._lib.uavnet_p2p_alloc(nbytes, C.byref(ptr), C.cast(hdl, vp))
            if rc:
                raise RuntimeError("uavnet_p2p_alloc failed (%d)" % rc)
            bufs.append(_ForeignCudaBuffer(ptr.value, nbytes // 4))
            handles.append(bytes(hdl))
        grad = torch.as_tensor(bufs[0], device=self.device)
        flat = torch.as_tensor(bufs[1], device=self.device)
        flat.copy_(self.flat)
        grad.copy_(self.grad)
        self.flat, self.grad = flat, grad
        self._bind_views()
        gptrs, pptrs, fptrs, opened = (vp * world)(), (vp * world)(), (vp * world)(), []
        gptrs[rank], pptrs[rank], fptrs[rank] = bufs[0].ptr, bufs[1].ptr, bufs[2].ptr
        if world > 1:
            every = [None] * world
            dist.all_gather_object(every, handles)
            for r in range(world):
                if r == rank:
                    continue
                for which, table in ((0, gptrs), (1, pptrs), (2, fptrs)):
                    peer = vp()
                    rc = self._lib.uavnet_p2p_open(C.cast(C.c_char_p(every[r][which]), vp), C.byref(peer))
                    if rc:
                        raise RuntimeError("uavnet_p2p_open failed (%d): no peer access to rank %d" % (rc, r))
                    table[r] = peer.value
                    opened.append(peer.value)
            torch.cuda.synchronize(self.device)
            dist.barrier()                            # every rank has mapped every buffer before anyone pushes
        # split: the push of an update comes in two parts -- the actor half of the first layer's gradient as soon as its gather
        # pass is done (push_early, on a side stream under the critic half's pass), everything else in apply_grads.  The two
        # modes give different ranks ownership of an element (its RMSProp slot lives on the owner), so the mode is fixed here.
        self._p2p = {"bufs": bufs, "gptrs": gptrs, "pptrs": pptrs, "fptrs": fptrs, "opened": opened, "rank": rank, "world": world,
                     "split": bool(split), "early_done": False}
        return self

    def _push_parts(self):
        """the two pushes of the split mode as lists of (offset, rows, row_width, col0, n_cols) in float32 elements"""
        o, n, stored, _ = self.segments["W1"]
        H2 = stored[1]
        early = [(o, stored[0], H2, 0, H2 // 2)]
        tail0 = o + (n + 3) // 4 * 4
        late = [(o, stored[0], H2, H2 // 2, H2 // 2), (tail0, 1, self.n_flat - tail0, 0, self.n_flat - tail0)]
        return early, late

    def _push_part(self, parts, lr: float):
        q = self._p2p
        a = N.PushPart(*parts[0])
        b = N.PushPart(*parts[1]) if len(parts) > 1 else None
        rc = self._lib.uavnet_p2p_push_part(q["gptrs"], q["pptrs"], q["fptrs"], _ptr(self.ms), C.byref(a), C.byref(b) if b is not None else None,
                                            q["rank"], q["world"], lr, RMS_DECAY, RMS_EPS, self._stream())
        if rc:
            raise RuntimeError("uavnet_p2p_push_part failed (%d)" % rc)

    def push_early(self, lr: float = LR_A):
        """split mode: push the actor half of the first layer's gradient now, on a side stream that waits for everything
        enqueued so far on the current one; apply_grads pushes the rest and joins"""
        q = self._p2p
        main = torch.cuda.current_stream(self.device)
        ps = self._side_stream(2)
        ps.wait_stream(main)
        with torch.cuda.stream(ps):
            self._push_part(self._push_parts()[0], lr)
        q["early_done"] = True

    def _apply_grads_p2p(self, lr: float):
        q = self._p2p
        if q["split"]:
            early, late = self._push_parts()
            if q["early_done"]:
                torch.cuda.current_stream(self.device).wait_stream(self._side_stream(2))    # pushes are ordered by their epochs
            else:
                self._push_part(early, lr)            # nobody pushed the first part ahead of time: both parts here
            self._push_part(late, lr)
            q["early_done"] = False
            return
        rc = self._lib.uavnet_p2p_push(q["gptrs"], q["pptrs"], q["fptrs"], _ptr(self.ms), self.n_flat, q["rank"], q["world"], lr,
                                       RMS_DECAY, RMS_EPS, self._stream())
        if rc:
            raise RuntimeError("uavnet_p2p_push failed (%d)" % rc)

    def p2p_status(self):
        """(pushes completed on this rank, True if a flag wait ever gave up) -- synchronises"""
        q = self._p2p
        torch.cuda.synchronize(self.device)
        e, t = C.c_uint32(0), C.c_uint32(0)
        rc = self._lib.uavnet_p2p_push_status(C.c_void_p(q["bufs"][2].ptr), C.byref(e), C.byref(t))
        if rc:
            raise RuntimeError("uavnet_p2p_push_status failed (%d)" % rc)
        return int(e.value), bool(t.value)

    def _p2p_owned_mask(self) -> torch.Tensor:
        """bool [n_flat]: the elements whose RMSProp slot this rank maintains (its slices of the pushes' element numberings,
        uavnet_p2p_push / uavnet_p2p_push_part)"""
        q = self._p2p
        world, rank = q["world"], q["rank"]
        pushes = list(self._push_parts()) if q["split"] else [[(0, 1, self.n_flat, 0, self.n_flat)]]
        mask4 = torch.zeros(self.n_flat // 4, dtype=torch.bool, device=self.device)
        for parts in pushes:
            sizes = [rows * (ncols // 4) for (_, rows, _, _, ncols) in parts]
            n4 = sum(sizes)
            per4 = (n4 + world - 1) // world
            lo, hi = min(rank * per4, n4), min(rank * per4 + per4, n4)
            j = torch.arange(lo, hi, device=self.device, dtype=torch.int64)
            start = 0
            for (off, rows, row_w, col0, ncols), sz in zip(parts, sizes):
                jj = j[(j >= start) & (j < start + sz)] - start
                sub = ncols // 4
                mask4[off // 4 + (jj // sub) * (row_w // 4) + col0 // 4 + jj % sub] = True
                start += sz
        return mask4.repeat_interleave(4)

    def close_p2p(self):
        """Back to private buffers.  The push keeps the RMSProp slot `ms` only for the slice a rank owns: the slices are
        gathered here so that every rank leaves with the complete optimiser state (and may continue with all-reduce +
        uavnet_rmsprop)."""
        q = getattr(self, "_p2p", None)
        if not q:
            return
        import torch.distributed as dist
        torch.cuda.synchronize(self.device)
        if q["world"] > 1:
            mine = torch.where(self._p2p_owned_mask(), self.ms, torch.zeros_like(self.ms))     # every element has one owner
            if dist.get_backend() == "nccl":
                dist.all_reduce(mine)
            else:                                     # gloo (tests): through host memory
                host = mine.cpu()
                dist.all_reduce(host)
                mine = host.to(self.device)
            self.ms.copy_(mine)
            torch.cuda.synchronize(self.device)
            dist.barrier()                            # nobody unmaps while a peer could still be in a push
        for ptr in q["opened"]:
            self._lib.uavnet_p2p_close(C.c_void_p(ptr))
        flat, grad = self.flat.clone(), self.grad.clone()
        self.flat, self.grad = flat, grad
        self._bind_views()
        for b in q["bufs"]:
            self._lib.uavnet_p2p_free(C.c_void_p(b.ptr))
        self._p2p = None
        self._pt_seen = None

    # ---- on-disk format of the reference (main.py:264-269,314; main_test.py:15-25) ---------------------------
    def actor_params(self):
        """[la/kernel, la/bias, la2/kernel, la2/bias, ap/kernel, ap/bias] as host arrays (GLOBAL_AC.a_params order)"""
        H, p = self.h, self.p
        return [p["W1"][:, :H].cpu().numpy().copy(), p["b1"][:H].cpu().numpy().copy(), p["Wa2"].cpu().numpy().copy(),
                p["ba2"].cpu().numpy().copy(), p["Wa3"].cpu().numpy().copy(), p["ba3"].cpu().numpy().copy()]

    def save_actor_npz(self, path: str):
        """np.savez(path, SESS.run(GLOBAL_AC.a_params)) (main.py:269): one ragged object array under 'arr_0'"""
        arr = np.empty(6, dtype=object)
        for i, a in enumerate(self.actor_params()):
            arr[i] = a
        np.savez(path, arr)

    def load_actor_npz(self, path: str):
        """main_test.py:15-25: assign arr_0[0..5] to la/kernel, la/bias, la2/kernel, la2/bias, ap/kernel, ap/bias"""
        a = np.load(path, allow_pickle=True)["arr_0"]
        H, p = self.h, self.p
        p["W1"][:, :H].copy_(torch.from_numpy(np.asarray(a[0], dtype=np.float32)))
        p["b1"][:H].copy_(torch.from_numpy(np.asarray(a[1], dtype=np.float32)))
        for k, i in (("Wa2", 2), ("ba2", 3), ("Wa3", 4), ("ba3", 5)):
            p[k].copy_(torch.from_numpy(np.asarray(a[i], dtype=np.float32)))


def n_step_targets(rewards: torch.Tensor, dones: torch.Tensor, v_boot: torch.Tensor, gamma: float = GAMMA,
                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Discounted n-step value targets of main.py:217-227, batched over envs: walking the buffer backwards,
    v = r + gamma * v, starting from the bootstrap value of the state after the last step -- 0 where the episode ended
    (`if done: v_s_ = 0`), and never carried across an episode end.  rewards/dones [T, E], v_boot [E] -> [T, E].
    float32 CUDA inputs run as one kernel (uavnet_nstep_targets); anything else (float64 checks, CPU) as torch ops."""
    T, E = rewards.shape
    if rewards.is_cuda and rewards.dtype == torch.float32 and v_boot.dtype == torch.float32 and dones.dtype in (torch.bool, torch.uint8) \
            and rewards.is_contiguous() and dones.is_contiguous() and v_boot.is_contiguous():
        if out is None:
            out = torch.empty_like(rewards)
        stream = C.c_void_p(torch.cuda.current_stream(rewards.device).cuda_stream)
        rc = N.lib().uavnet_nstep_targets(_ptr(rewards), _ptr(dones), _ptr(v_boot), T, E, float(gamma), _ptr(out), stream)
        if rc:
            raise RuntimeError("uavnet_nstep_targets failed (%d)" % rc)
        return out
    out = torch.empty_like(rewards)
    v = v_boot
    for t in range(T - 1, -1, -1):
        v = rewards[t] + gamma * torch.where(dones[t].bool(), torch.zeros_like(v), v)
        out[t] = v
    return out


class A3CTrainer:
    """Synchronous batched restatement of Worker.work (main.py:182-271) for one rank: E envs, rollouts of
    UPDATE_GLOBAL_ITER steps, one update per rollout.  With torch.distributed initialised the gradient buffer is
    all-reduced (NCCL over NVLink) before the optimiser pass -- the reference's push/pull (main.py:159-163).

    ``env`` may be a list of BatchedMobiEnvironment handles with consecutive ``env_offset`` ranges (the reference's
    independent workers, main.py:173): every handle rolls out on its own CUDA stream, so the short kernels of one
    group's step (policy forward of 64 CTAs, the latency-bound env step) overlap the other groups'.  All draws are keyed
    by the global env id, so the rollout is the same for any grouping."""

    def __init__(self, env, net: ACNet, rollout: int = UPDATE_GLOBAL_ITER, seed: int = 0):
        self.envs = list(env) if isinstance(env, (list, tuple)) else [env]
        self.env = self.envs[0]
        self.net, self.T = net, int(rollout)
        self.K = self.env.nUE + self.env.nBS
        dev = self.env.device
        self.slices, lo = [], 0
        for i, e in enumerate(self.envs):
            if e.device != dev or e.nUE + e.nBS != self.K:
                raise ValueError("all env handles must live on one device and have the same shape")
            if i and e.env_offset != self.envs[i - 1].env_offset + self.envs[i - 1].n_envs:
                raise ValueError("env handles must cover consecutive global env ranges (env_offset)")
            self.slices.append(slice(lo, lo + e.n_envs))
            lo += e.n_envs
        self.E = lo
        self.streams = [torch.cuda.Stream(device=dev) for _ in self.envs] if len(self.envs) > 1 else [None]
        self.seed = int(seed)
        self._draws = torch.zeros(1, dtype=torch.int32, device=dev)          # rollout steps sampled so far (Philox counter)
        # T + 1 slots: the env writes the state after step t straight into slot t + 1 (bind_obs_idx); slot T is the
        # bootstrap state and becomes slot 0 of the next rollout
        self.buf_idx = torch.empty((self.T + 1, self.E, self.K), dtype=torch.int32, device=dev)
        self.buf_vt = torch.empty((self.T, self.E), dtype=torch.float32, device=dev)
        self.buf_a = torch.empty((self.T, self.E), dtype=torch.int64, device=dev)
        self.buf_r = torch.empty((self.T, self.E), dtype=torch.float32, device=dev)
        self.buf_done = torch.empty((self.T, self.E), dtype=torch.bool, device=dev)
        # activations of the rollout's forward passes, reused by the update (same parameters)
        self.buf_h1 = torch.empty((self.T, self.E, 2 * net.h), dtype=torch.float32, device=dev)
        self.buf_h2a = torch.empty((self.T, self.E, net.h), dtype=torch.float32, device=dev)
        self.buf_prob = torch.empty((self.T, self.E, net.n_a), dtype=torch.float32, device=dev)
        self.buf_logits = torch.empty((self.E, net.n_a), dtype=torch.float32, device=dev)
        self.ep_return = torch.zeros(self.E, dtype=torch.float64, device=dev)       # ep_r of the running episodes
        self.ep_finished = torch.full((self.E,), float("nan"), dtype=torch.float64, device=dev)   # return of each env's last finished episode
        for e, sl in zip(self.envs, self.slices):
            e.bind_obs_idx(self.buf_idx[0][sl])
            e.reset()
        # host-side lower bound on the steps until an env of handle g can be done (all envs were just reset)
        self._until = [int(e.cfg.max_step) for e in self.envs]
        self._stale = [False] * len(self.envs)
        self._resets = None
        self.updates = 0

    def _rollout_step(self, g: int, t: int):
        env, net, sl = self.envs[g], self.net, self.slices[g]
        h1 = net.first_layer(self.buf_idx[t][sl], self.buf_h1[t][sl])    # s_t, written there by the env itself
        h2a = net.actor_hidden(h1, self.buf_h2a[t][sl])
        # softmax + np.random.choice(p=a_prob) (main.py:149,165-169,195) fused; draws keyed by the GLOBAL env id
        _, a = net.sample_head(h2a, self.seed, env.env_offset, self._draws, t, prob_out=self.buf_prob[t][sl],
                               action_out=self.buf_a[t][sl], logits_out=self.buf_logits[sl])
        env.bind_obs_idx(self.buf_idx[t + 1][sl])
        _, r, _, _ = env.step(a)                                         # main.py:198
        rc = net._lib.uavnet_rollout_record(_ptr(r), _ptr(env.done_u8), env.n_envs, _ptr(self.buf_r[t][sl]),
                                            _ptr(self.buf_done[t][sl]), _ptr(self.ep_return[sl]), _ptr(self.ep_finished[sl]),
                                            net._stream())
        if rc:
            raise RuntimeError("uavnet_rollout_record failed (%d)" % rc)
        # finished episodes restart (main.py:188-190).  The masked reset is a launch of E CTAs that all exit unless some env
        # is done; the host knows a lower bound on the steps until that can happen (`_until`), so the launch is only made
        # when an episode CAN end at this step -- 1 rollout in 200 at the reference's MAXSTEP = 2000, T = 10.
        self._until[g] -= 1
        if self._resets is True or (self._resets is None and self._until[g] <= 0):
            env.reset(env_mask=env.done_u8)
            self._stale[g] = True

    def _resync_until(self):
        """after a rollout that may have reset some envs: steps until the earliest possible `done` per handle (one small
        device read per episode end)"""
        for g, e in enumerate(self.envs):
            if self._stale[g]:
                self._until[g] = int((int(e.cfg.max_step) - e.step_n).min())
                self._stale[g] = False

    def rollout(self, resets=None):
        """resets: None = the masked reset is launched only at steps where an episode can end (host-side bound); True /
        False = always / never (what `capture` records into its two graphs)"""
        net = self.net
        self._resets = resets
        net._sync_transposed()                                               # before the streams fork
        if len(self.envs) == 1:
            for t in range(self.T):
                self._rollout_step(0, t)
        else:
            cur = torch.cuda.current_stream(self.env.device)
            for st in self.streams:
                st.wait_stream(cur)
            for t in range(self.T):                                          # issue order interleaves the groups
                for g, st in enumerate(self.streams):
                    with torch.cuda.stream(st):
                        self._rollout_step(g, t)
            for st in self.streams:
                cur.wait_stream(st)
        self._draws += self.T                                                # next rollout: fresh Philox counters
        v_boot = net.value(self.buf_idx[self.T])                             # main.py:217-220
        vt = n_step_targets(self.buf_r, self.buf_done, v_boot, out=self.buf_vt)
        self.buf_idx[0].copy_(self.buf_idx[self.T])                          # the next rollout starts where this one ended
        for e, sl in zip(self.envs, self.slices):
            e.bind_obs_idx(self.buf_idx[0][sl])
        if resets is None and any(self._stale):
            self._resync_until()
        return vt

    def update(self, v_target: torch.Tensor):
        net, M = self.net, self.T * self.E
        saved = {"h1": self.buf_h1.view(M, -1), "h2a": self.buf_h2a.view(M, -1), "prob": self.buf_prob.view(M, -1)}
        q = getattr(net, "_p2p", None)
        early = net.push_early if (q and q["split"] and q["world"] > 1 and net.sparse_bwd == "gather") else None
        a_loss, c_loss = net.accumulate_grads(self.buf_idx[:self.T].view(M, self.K), self.buf_a.view(M), v_target.reshape(M), saved,
                                              early_push=early)
        world = 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            world = torch.distributed.get_world_size()
            if world > 1 and not getattr(net, "_p2p", None):
                torch.distributed.all_reduce(net.grad)                       # the gradient push, summed over ranks
        net.apply_grads(LR_A, world)
        self.updates += 1
        return a_loss, c_loss

    def train_iteration(self, resets=None):
        return self.update(self.rollout(resets))

    # ---- one CUDA graph per iteration: ~100 launches (env step, reset, dense products, softmax + sampling, ...) replayed
    # ---- without Python or launch latency between them
    def capture(self, warmup: int = 2):
        """Capture ``train_iteration`` into a CUDA graph (after `warmup` eager iterations on a side stream).
        The env / net / trainer buffers are all persistent, so the captured pointers stay valid; the sampling kernel
        reads its Philox draw counter from device memory (``self._draws``, advanced inside the graph), so every replay
        draws fresh actions.  The stream forks of a grouped rollout are captured with it."""
        dev = self.env.device
        warmup = max(1, int(warmup))      # first launches set function attributes / allocate the error word: not capturable
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.train_iteration()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        # two graphs: the iteration with the masked resets after every step and the one without; the host picks per replay
        # from its bound on the steps until an episode can end
        until = list(self._until)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._graph_out = self.train_iteration(resets=True)
        self._graph_lean = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph_lean):
            self._graph_out_lean = self.train_iteration(resets=False)
        self._until = until                       # capturing runs nothing: the counters go back to where they were
        self._stale = [False] * len(self.envs)
        self._resets = None
        return self

    def train_iteration_graph(self):
        """Replay the captured iteration; returns the (a_loss, c_loss) tensors of the replay (static buffers)."""
        if min(self._until) > self.T:             # no episode can end inside this rollout: the graph without reset launches
            self._until = [u - self.T for u in self._until]
            self._graph_lean.replay()
            return self._graph_out_lean
        self._graph.replay()
        self._stale = [True] * len(self.envs)
        self._resync_until()
        return self._graph_out
