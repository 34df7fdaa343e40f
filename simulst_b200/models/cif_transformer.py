"""Drop-in bodies for ``CIFLayer.forward`` / ``CIFLayer.infer`` of the reference
(codebase/models/cif_transformer.py:143-261).  The alpha projection
(CausalConvTBC -> LayerNorm -> GELU -> Dropout -> Linear, :124-130) stays the reference's;
``B200CIFLayerMixin`` goes in front of the reference class and only swaps ``cif_function``."""
from typing import Dict, Optional

import torch
from torch import Tensor

from .torch_cif import cif_function


class B200CIFLayerMixin:
    """Expects the attributes of the reference CIFLayer: alpha_proj, sg_alpha, beta, tail_thres,
    get_incremental_state / set_incremental_state."""

    def _integration_weights(self, x, incremental_state=None):
        """reference :158-161 (train) / :202-208 (infer): sigmoid(alpha_proj(x)) -> (B, S)."""
        if incremental_state is None:
            alpha = self.alpha_proj(x.detach() if self.sg_alpha else x)
        else:
            alpha = x
            for m in self.alpha_proj:
                try:
                    alpha = m(alpha, incremental_state)
                except TypeError:
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
        alpha = self._integration_weights(x, incremental_state if incremental_state is not None else {})
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
    Mutates `cached_state` (prev_weight (B,1), prev_feat (B,1,C))."""
    bsz = x.size(1)
    x = x.transpose(1, 0)
    if (
        "prev_weight" in cached_state
        and cached_state["prev_weight"] is not None
        and cached_state["prev_weight"].numel() > 0
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
