"""Fused hard-negative-mining InfoNCE on B200 (addition next to the reference's `atq` names).

`HardNegativeMiningInfoNCE` mirrors utils/enhanced_contrastive.py:8-162 -- same constructor, `set_epoch`,
`get_current_temperature`, `forward(image_embeddings, text_embeddings, weights=None)` -- and evaluates the loss with
the kernels of csrc/loss_sm100.cu: the B x B similarity matrix comes from the ternary path's tcgen05 GEMM (scaled
fp16 operand pairs, 1/temperature folded into the epilogue), the two top-k hardness masks become per-row / per-column
k-th-largest thresholds (exact radix select), cross entropy + entropy regulariser are two passes of per-row
statistics, and backward is one elementwise kernel over the matrix plus two GEMMs.  The reference builds the masks
with a Python loop of B indexed writes (:118-120) and ~25 full-matrix torch ops.
`ContrastiveLearningManager` mirrors :269-417 (curriculum weights; the weights keep their gradient, as there).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch.nn.modules.loss import _Loss

from . import _engine as eng
from . import _native as nv


def _vec(n, dev):
    return torch.empty(n, dtype=torch.float32, device=dev)


class _InfoNCEFn(torch.autograd.Function):
    """loss(img_n, txt_n, pos_weights) for L2-normalised [B, E] embeddings."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, img, txt, pw, temperature, lambda_reg, hard_mul, k):
        img = nv.require_f32(img, "image_embeddings")
        txt = nv.require_f32(txt, "text_embeddings")
        B, E = img.shape
        dev = nv.device_index(img)
        d = img.device
        st = nv.stream_ptr(dev)
        inv_t = torch.full((1,), 1.0 / float(temperature), dtype=torch.float32, device=d)
        ia, ta = eng.split_operand(img), eng.split_operand(txt)
        S, _ = eng.tgemm(ia, ta, B, B, E, scale=inv_t)           # sim[i, j] / temperature
        St = S.t().contiguous()
        thr_r, thr_c = _vec(B, d), _vec(B, d)
        if B >= 2 and k <= B - 1:
            nv.call("atq_rowkth_largest", dev, S.data_ptr(), B, B, int(k), thr_r.data_ptr(), st)
            nv.call("atq_rowkth_largest", dev, St.data_ptr(), B, B, int(k), thr_c.data_ptr(), st)
        else:  # topk(k = B) keeps every entry: every negative is "hard"
            thr_r.fill_(-math.inf)
            thr_c.fill_(-math.inf)
        pwc = None if pw is None else nv.require_f32(pw.detach().reshape(-1), "weights")
        lw_r, ls_r, es_r, wd = _vec(B, d), _vec(B, d), _vec(B, d), _vec(B, d)
        lw_c, ls_c, es_c = _vec(B, d), _vec(B, d), _vec(B, d)
        nv.call("atq_infonce_row_stats", dev, S.data_ptr(), B, B, thr_r.data_ptr(), thr_c.data_ptr(), nv.ptr(pwc), float(hard_mul),
                lw_r.data_ptr(), ls_r.data_ptr(), es_r.data_ptr(), wd.data_ptr(), st)
        nv.call("atq_infonce_row_stats", dev, St.data_ptr(), B, B, thr_c.data_ptr(), thr_r.data_ptr(), nv.ptr(pwc), float(hard_mul),
                lw_c.data_ptr(), ls_c.data_ptr(), es_c.data_ptr(), None, st)
        loss = torch.empty((), dtype=torch.float32, device=d)
        nv.call("atq_infonce_finalize", dev, B, lw_r.data_ptr(), lw_c.data_ptr(), ls_r.data_ptr(), ls_c.data_ptr(), es_r.data_ptr(),
                es_c.data_ptr(), wd.data_ptr(), float(lambda_reg), loss.data_ptr(), st)
        ctx.save_for_backward(S, thr_r, thr_c, lw_r, lw_c, ls_r, ls_c, es_r, es_c, *( [pwc] if pwc is not None else []))
        ctx.ops = (ia, ta)
        ctx.cfg = (B, E, float(temperature), float(lambda_reg), float(hard_mul), pw is not None)
        return loss

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, go):
        B, E, temperature, lambda_reg, hard_mul, has_pw = ctx.cfg
        saved = ctx.saved_tensors
        S, thr_r, thr_c, lw_r, lw_c, ls_r, ls_c, es_r, es_c = saved[:9]
        pwc = saved[9] if has_pw else None
        ia, ta = ctx.ops
        dev = nv.device_index(S)
        d = S.device
        go = nv.require_f32(go.reshape(1), "grad_output")
        dS = torch.empty_like(S)
        dpw = _vec(B, d) if has_pw else None
        nv.call("atq_infonce_grad", dev, S.data_ptr(), B, B, thr_r.data_ptr(), thr_c.data_ptr(), nv.ptr(pwc), hard_mul,
                lw_r.data_ptr(), lw_c.data_ptr(), ls_r.data_ptr(), ls_c.data_ptr(), es_r.data_ptr(), es_c.data_ptr(), lambda_reg,
                1.0 / temperature, go.data_ptr(), dS.data_ptr(), B, nv.ptr(dpw), nv.stream_ptr(dev))
        ga = eng.split_operand(dS)
        # d img = dS . txt (txt [B, E] consumed MN-major);  d txt = dS^T . img (both operands MN-major)
        d_img, _ = eng.tgemm(ga, eng.mn_view(ta), B, E, B)
        d_txt, _ = eng.tgemm_dw_masked(eng.mn_view(ga), eng.mn_view(ia), B, E, B)
        return d_img, d_txt, dpw, None, None, None, None


def hard_negative_infonce(image_embeddings, text_embeddings, weights=None, temperature=0.07, lambda_reg=0.02,
                          hard_negative_weight=0.5, hardest_mining_ratio=0.5):
    """The loss of utils/enhanced_contrastive.py:64-158 for the given temperature (fused kernels, CUDA only)."""
    img = F.normalize(image_embeddings, p=2, dim=1)
    txt = F.normalize(text_embeddings, p=2, dim=1)
    b = img.size(0)
    k = max(1, int(b * hardest_mining_ratio))
    return _InfoNCEFn.apply(img, txt, weights, temperature, lambda_reg, 1.0 + hard_negative_weight, k)


class HardNegativeMiningInfoNCE(_Loss):
    """Drop-in for utils/enhanced_contrastive.py:8 (same arguments, same schedule, fused evaluation)."""

    def __init__(self, temperature=0.07, lambda_reg=0.02, hard_negative_weight=0.5, hardest_mining_ratio=0.5,
                 temperature_schedule=True):
        super().__init__()
        self.temperature = temperature
        self.lambda_reg = lambda_reg
        self.hard_negative_weight = hard_negative_weight
        self.hardest_mining_ratio = hardest_mining_ratio
        self.temperature_schedule = temperature_schedule
        self.base_temperature = temperature
        self.current_epoch = 0
        self.total_epochs = 1

    def set_epoch(self, current_epoch, total_epochs):
        self.current_epoch = current_epoch
        self.total_epochs = total_epochs

    def get_current_temperature(self):
        # utils/enhanced_contrastive.py:47-62 (cosine annealing from 2x to 0.5x the base temperature)
        if not self.temperature_schedule:
            return self.temperature
        progress = min(1.0, self.current_epoch / (self.total_epochs * 0.7))
        hi, lo = self.base_temperature * 2.0, self.base_temperature * 0.5
        t = hi - (hi - lo) * (1 - math.cos(progress * math.pi)) / 2
        return max(min(t, hi), lo)

    def forward(self, image_embeddings, text_embeddings, weights=None):
        return hard_negative_infonce(image_embeddings, text_embeddings, weights, self.get_current_temperature(), self.lambda_reg,
                                     self.hard_negative_weight, self.hardest_mining_ratio)


class ContrastiveLearningManager:
    """utils/enhanced_contrastive.py:269-417: curriculum stage from the epoch, curriculum weights from the positives'
    cosine similarity (only the diagonal is evaluated here -- the reference forms the whole B x B product for it)."""

    def __init__(self, model=None, criterion=None, similarity_threshold=0.8, mining_freq=50, curriculum_steps=3):
        self.model, self.criterion = model, criterion
        self.similarity_threshold, self.mining_freq, self.curriculum_steps = similarity_threshold, mining_freq, curriculum_steps
        self.steps = self.epoch = self.total_epochs = self.curriculum_stage = 0
        self.mined_examples = []

    def set_epoch(self, epoch, total_epochs):
        self.epoch, self.total_epochs = epoch, total_epochs
        self.curriculum_stage = min(self.curriculum_steps - 1, int(epoch / total_epochs * self.curriculum_steps))

    def get_curriculum_weight(self, similarity=None, positives=None):
        pos = torch.diag(similarity) if positives is None else positives
        if self.curriculum_stage == 0:
            return torch.sigmoid(pos * 10)
        if self.curriculum_stage == self.curriculum_steps - 1:
            return 1 - torch.sigmoid(pos * 10 - 5)
        return torch.ones_like(pos)

    def compute_loss(self, image_embeddings, text_embeddings, similarity=None):
        self.steps += 1
        if similarity is None:
            pos = (F.normalize(image_embeddings, p=2, dim=1) * F.normalize(text_embeddings, p=2, dim=1)).sum(dim=1)
            weights = self.get_curriculum_weight(positives=pos)
        else:
            weights = self.get_curriculum_weight(similarity)
        return self.criterion(image_embeddings, text_embeddings, weights)
