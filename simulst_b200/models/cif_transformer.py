"""Drop-in bodies for ``CIFLayer.forward`` / ``CIFLayer.infer`` of the reference
(codebase/models/cif_transformer.py:143-261).  The alpha projection
(CausalConvTBC -> LayerNorm -> GELU -> Dropout -> Linear, :124-130) stays the reference's;
``B200CIFLayerMixin`` goes in front of the reference class and only swaps ``cif_function``."""
from typing import Dict, Optional

import torch
from torch import Tensor

from .torch_cif import cif_function


def _takes_incremental_state(module) -> bool:
    """CausalConvTBC (modules/causal_conv.py:80-98) is ``make_causal(ConvTBC)`` decorated with
    ``with_incremental_state``; LayerNorm / GELU / Dropout / Linear are not."""
    return type(module).__name__ == "CausalConvTBC" or hasattr(module, "get_incremental_state")


class B200CIFLayerMixin:
    """Expects the attributes of the reference CIFLayer: alpha_proj, sg_alpha, beta, tail_thres,
    get_incremental_state / set_incremental_state."""

    def _integration_weights(self, x, incremental_state=None, streaming=False):
        """reference :157-161 (train) / :201-208 (infer): sigmoid(alpha_proj(x)) -> (B, S).
        In streaming mode the reference hands ``incremental_state`` (unchanged, None included) to
        the ``CausalConvTBC`` members of ``alpha_proj`` and calls the others plainly; the members
        are recognised by what makes them streamable -- the incremental-state accessors that
        fairseq's ``with_incremental_state`` puts on the class -- not by catching TypeError."""
        if not streaming:
            alpha = self.alpha_proj(x.detach() if self.sg_alpha else x)
        else:
            alpha = x
            for m in self.alpha_proj:
                if _takes_incremental_state(m):
                    alpha = m(alpha, incremental_state)
                else:
                    alpha = m(alpha)
        return alpha.transpose(1, 0).sigmoid().squeeze(-1)

    def forward(
        self,
        x,
        encoder_padding_mask: Optional[Tensor] = None,
        target_lengths: Optional[Tensor] = None,
    ):
        """reference :143-186.  x: (seq_len, batch, embed_dim)."""
        alpha = self._integration_weights(x)
        return cif_layer_forward(x, alpha, self.beta, self.tail_thres, encoder_padding_mask,
                                 target_lengths)

    def infer(
        self,
        x,
        incremental_state: Optional[Dict[str, Dict[str, Optional[Tensor]]]] = None,
        encoder_padding_mask: Optional[Tensor] = None,
        finish=False,
    ):
        """reference :188-261.  x: (chunk_len, 1, embed_dim), a new chunk."""
        chunk_len, bsz, C = x.size()
        if bsz > 1:
            raise NotImplementedError("batched infer not supported for now.")
        alpha = self._integration_weights(x, incremental_state, streaming=True)
        cached_state = self.get_incremental_state(incremental_state, "cif_state")
        if cached_state is None:
            cached_state = {}
        out = cif_layer_infer(x, alpha, cached_state, self.beta, self.tail_thres, finish)
        self.set_incremental_state(incremental_state, "cif_state", cached_state)
        return out


def cif_layer_forward(x: Tensor, alpha: Tensor, beta: float, tail_thres: float,
                      encoder_padding_mask: Optional[Tensor] = None,
                      target_lengths: Optional[Tensor] = None):
    """reference :163-186 from the point where `alpha` is the (B, S) sigmoid output."""
    x = x.transpose(1, 0)
    # apply masking first
    if encoder_padding_mask is not None:
        x = x.masked_fill(encoder_padding_mask.unsqueeze(2), 0)
        alpha = alpha.masked_fill(encoder_padding_mask, 0)
    cif_out = cif_function(
        x,
        alpha,
        beta=beta,
        tail_thres=tail_thres,
        target_lengths=target_lengths,
    )
    # (B, T, C) -> (T, B, C)
    cif_feats = cif_out["cif_out"][0].transpose(0, 1)
    cif_out.update({
        "cif_out": [cif_feats],
        "alpha": [alpha]
    })
    return cif_out


def cif_layer_infer(x: Tensor, alpha: Tensor, cached_state: Dict[str, Optional[Tensor]],
                    beta: float, tail_thres: float, finish: bool = False):
    """reference :210-261 from the point where `alpha` is the (1, chunk) sigmoid output.
    Mutates `cached_state` (prev_weight (B,1), prev_feat (B,1,C)).

    Upstream quirk, REPRODUCED: between chunks ``tail_thres`` is 0, so the
    tail always "fires" and is rescaled by ``beta / tail_weight`` (cif.py:170-178); a chunk whose
    accumulated weight ends exactly on a firing boundary (``tail_weight == 0``) therefore carries
    ``0 * inf = NaN`` features with weight 0 into the next chunk, in the reference and here alike
    (``tests/test_reference_classes_gpu.py::test_cif_infer_zero_tail_weight_quirk``).  Guarding it
    would change results a fairseq checkpoint was trained against; a caller that wants the guard
    can zero ``prev_feat`` where ``prev_weight == 0``."""
    bsz = x.size(1)
    x = x.transpose(1, 0)
    if (
        "prev_weight" in cached_state
        and cached_state["prev_weight"].numel() > 0     # None after finish=True: raises, as upstream (:254)
    ):
        # leftover features with weight: treated as a single source feature
        alpha = torch.cat((cached_state["prev_weight"], alpha), dim=1)
        x = torch.cat((cached_state["prev_feat"], x), dim=1)

    cif_out = cif_function(
        x,
        alpha,
        beta=beta,
        tail_thres=tail_thres if finish else 0,
    )
    cif_feats = cif_out["cif_out"][0]  # (B, t, C)
    cif_len = cif_out["cif_lengths"][0]  # (B,)
    tail_weight = cif_out["tail_weights"][0]  # (B,)
    n_fired = int(cif_len.item())            # B = 1 (reference :239)
    if not finish:
        prev_feat = cif_feats[:, n_fired - 1:, :]  # (B, 1, C)
        prev_weight = tail_weight.view(bsz, 1)   # (B, 1)
        # feat was normalized to beta in cif_function(), unscale to 1 for next segment
        prev_feat = prev_feat / beta
    else:
        prev_feat = None
        prev_weight = None
    cached_state["prev_feat"] = prev_feat
    cached_state["prev_weight"] = prev_weight

    cif_len = cif_len if finish else (cif_len - 1)
    cif_feats = cif_feats.narrow(
        1, 0, n_fired if finish else n_fired - 1).transpose(0, 1)  # (B, t-1, C)
    cif_out.update({
        "cif_out": [cif_feats],
        "cif_lengths": [cif_len],
        "alpha": [alpha]
    })
    return cif_out
