"""Generator composition with the reference's forward signature (graph/model.py:12-41).

``Model.forward(note, pre_note, phrase, position, is_train=True)`` returns ``(gen_note, z, pre_z, phrase_feature)``
in training mode and the generated bar when ``is_train=False`` (sampling: the ``note`` slot carries the latent,
maker_bar.py:38).  The reference's Refiner cannot execute (graph/refiner.py:12 vs :19), so --
exactly like the reference's own runnable composition graph/model_with_gan.py:20-38 -- it is not applied here.

``vae_head=True`` adds the reparameterise + KL head of the archived generation
(old/graphs/models/bar_v1/encoder.py:55-63): forward then returns ``(recon, mu, logvar)``.
"""
import os

import torch
import torch.nn as nn

from .. import _lib
from ..engine import flatten
from .decoder import Decoder
from .encoder import Encoder
from .phrase_encoder import PhraseModel
from .weights_initializer import weights_init


class _ReparamFn(torch.autograd.Function):
    """z = mu + eps * exp(0.5 * logvar), one fused kernel each way (KL gradient is added by VAELoss)."""

    @staticmethod
    def forward(ctx, mu, logvar, eps):
        mu, logvar, eps = mu.contiguous(), logvar.contiguous(), eps.contiguous()
        z = torch.empty_like(mu)
        kl = torch.zeros(1, device=mu.device)
        _lib.check(_lib.lib().bvae_reparam_kl_fwd(mu.data_ptr(), logvar.data_ptr(), eps.data_ptr(), z.data_ptr(),
                                                  kl.data_ptr(), mu.numel(), _lib.stream_ptr()), "reparam_kl_fwd")
        ctx.save_for_backward(mu, logvar, eps)
        ctx.mark_non_differentiable(kl)
        return z, kl

    @staticmethod
    def backward(ctx, dz, _dkl):
        mu, logvar, eps = ctx.saved_tensors
        dz = dz.contiguous()
        dmu, dlv = torch.empty_like(mu), torch.empty_like(mu)
        _lib.check(_lib.lib().bvae_reparam_kl_bwd(mu.data_ptr(), logvar.data_ptr(), eps.data_ptr(), dz.data_ptr(), 0.0,
                                                  dmu.data_ptr(), dlv.data_ptr(), mu.numel(), _lib.stream_ptr()),
                   "reparam_kl_bwd")
        return dmu, dlv, None


class _KLFn(torch.autograd.Function):
    """-0.5 * sum(1 + logvar - mu^2 - exp(logvar))  (old/graphs/losses/loss.py:16)."""

    @staticmethod
    def forward(ctx, mu, logvar):
        mu, logvar = mu.contiguous(), logvar.contiguous()
        z = torch.empty_like(mu)
        kl = torch.zeros(1, device=mu.device)
        zero = torch.zeros_like(mu)
        _lib.check(_lib.lib().bvae_reparam_kl_fwd(mu.data_ptr(), logvar.data_ptr(), zero.data_ptr(), z.data_ptr(),
                                                  kl.data_ptr(), mu.numel(), _lib.stream_ptr()), "reparam_kl_fwd")
        ctx.save_for_backward(mu, logvar)
        return kl[0]

    @staticmethod
    def backward(ctx, g):
        mu, logvar = ctx.saved_tensors
        dmu, dlv = torch.empty_like(mu), torch.empty_like(mu)
        _lib.check(_lib.lib().bvae_reparam_kl_bwd(mu.data_ptr(), logvar.data_ptr(), mu.data_ptr(), None, 1.0,
                                                  dmu.data_ptr(), dlv.data_ptr(), mu.numel(), _lib.stream_ptr()),
                   "reparam_kl_bwd")
        return dmu * g, dlv * g


def reparameterize(mu, logvar, eps=None):
    """old/graphs/models/bar_v1/encoder.py:60-63; eps defaults to torch.randn_like (CUDA Philox)."""
    if eps is None:
        eps = torch.randn_like(mu)
    return _ReparamFn.apply(mu, logvar, eps)[0]


def kl_divergence(mu, logvar):
    return _KLFn.apply(mu, logvar)


def _cat_bars(note, pre_note):
    """torch.cat((note, pre_note), 0) -- as a VIEW when the two already sit back to back in one buffer (the bit-packed
    input path expands both into one [2B,1,96,60] tensor, data/packed.py)."""
    if (note.dtype == pre_note.dtype and note.shape == pre_note.shape and note.is_contiguous()
            and pre_note.is_contiguous() and not note.requires_grad and not pre_note.requires_grad
            and note.untyped_storage().data_ptr() == pre_note.untyped_storage().data_ptr()
            and pre_note.storage_offset() == note.storage_offset() + note.numel()):
        shape = (2 * note.shape[0],) + tuple(note.shape[1:])
        return note.as_strided(shape, torch.empty(shape, device="meta").stride(), note.storage_offset())
    return torch.cat((note, pre_note), 0)


class Model(nn.Module):
    def __init__(self, vae_head: bool = False, refiner: bool = False):
        super().__init__()
        self.encoder = Encoder([64, 128, 256, 512, 1024])
        self.decoder = Decoder([1024, 512, 256, 128, 64])
        self.phrase_encoder = PhraseModel([64, 128, 256, 512, 1024])
        # graph/model.py:18,31,41 applies a Refiner to the generated bar in both branches, but the reference's Refiner
        # cannot execute (graph/refiner.py:12 vs :19); ``refiner=True`` adds it WITH the one-line shape fix (see
        # graph/refiner.py here).  Off by default == the reference's runnable composition graph/model_with_gan.py.
        self.refiner = None
        if refiner:
            from .refiner import Refiner
            self.refiner = Refiner()
        self.vae_head = vae_head
        if vae_head:
            # log-variance head next to the mean head (= encoder.linear), as `var` sits next to `mean` in
            # old/graphs/models/bar_v1/encoder.py:55-56.  Not part of the HEAD state_dict.
            self.logvar_head = nn.Linear(1152, 1152)
            nn.init.zeros_(self.logvar_head.weight)
            nn.init.constant_(self.logvar_head.bias, -4.0)
        self.apply(weights_init) if not vae_head else [m.apply(weights_init) for m in
                                                        (self.encoder, self.decoder, self.phrase_encoder)]

    def flatten_parameters(self):
        """Re-home all parameters into one flat fp32 bucket (engine.FlatParams) -- used by the fused Adam step
        and the NCCL gradient all-reduce.  state_dict()/load_state_dict() keep working (parameters become views)."""
        return flatten(self)

    def side_stream(self, device):
        """the stream the phrase encoder runs on (None when BVAE_STREAMS=0)"""
        if os.environ.get("BVAE_STREAMS", "1") == "0":
            return None
        side = getattr(self, "_side_stream", None)
        if side is None or side.device != torch.device(device):
            side = self._side_stream = torch.cuda.Stream(device=device)
        return side

    def _phrase_branch(self, phrase):
        """The phrase encoder does not depend on the bar encoder: run it on a second stream so that its latency-bound
        small-map kernels and every kernel's tail overlap the other branch (autograd runs the backward node on the
        same stream; _TrunkFn.backward makes the caller's stream wait for it).  BVAE_STREAMS=0 disables."""
        if os.environ.get("BVAE_STREAMS", "1") == "0" or not phrase.is_cuda:
            return self.phrase_encoder(phrase)
        main = torch.cuda.current_stream()
        side = self.side_stream(phrase.device)
        side.wait_stream(main)
        if not torch.cuda.is_current_stream_capturing():      # (graph capture: the inputs are the graph's static buffers)
            phrase.record_stream(side)  # the caller may drop its input right after the call (one block per step)
        with torch.cuda.stream(side):
            pf = self.phrase_encoder(phrase)
        self._phrase_join = (main, side, pf)
        return pf

    def _join_phrase(self):
        j = getattr(self, "_phrase_join", None)
        if j is not None:
            main, side, pf = j
            # no record_stream on the phrase feature: it is saved by the decoder node, and the side stream next touches
            # its block only after side.wait_stream(main) of the following step (record_stream-ing every tensor that
            # crosses streams made the caching allocator pile up deferred frees and fall back to cudaMalloc)
            main.wait_stream(side)
            self._phrase_join = None

    def forward(self, note, pre_note, phrase, position, is_train=True, dropout_masks=None, eps=None):
        if is_train:
            B = note.shape[0]
            phrase_feature = self._phrase_branch(phrase)
            # encoder(note) and encoder(pre_note) share weights and have no batch-coupled op (InstanceNorm is
            # per sample): one pass over 2B bars (model.py:26-27)
            zz = self.encoder(_cat_bars(note, pre_note))
            z, pre_z = zz[:B], zz[B:]
            if self.vae_head:
                mu, pre_mu = z, pre_z
                lv_all = self.logvar_head(zz)
                logvar, pre_logvar = lv_all[:B], lv_all[B:]
                e1, e2 = (None, None) if eps is None else eps
                z = reparameterize(mu, logvar, e1)
                pre_z = reparameterize(pre_mu, pre_logvar, e2)
                self._join_phrase()
                recon = self.decoder(z, pre_z, phrase_feature, position, dropout_masks)
                self.last_pre = (pre_mu, pre_logvar)
                return recon, mu, logvar
            self._join_phrase()
            gen_note = self.decoder(z, pre_z, phrase_feature, position, dropout_masks)
            if self.refiner is not None:
                gen_note = self.refiner(gen_note)                       # graph/model.py:31
            return gen_note, z, pre_z, phrase_feature
        phrase_feature = self._phrase_branch(phrase)
        pre_z = self.encoder(pre_note)
        self._join_phrase()
        gen_note = self.decoder(note, pre_z, phrase_feature, position, dropout_masks)
        return gen_note if self.refiner is None else self.refiner(gen_note)   # graph/model.py:41
