"""Mirror of the reference's encoder wrapper (avhubert/hubert_asr.py:375-409): the object MMS-LLaMA owns as
``self.avhubert`` (src/model.py:96-97,220) and calls as ``self.avhubert(source={...}, padding_mask=...)``."""
import torch.nn as nn


class HubertEncoderWrapper(nn.Module):
    def __init__(self, w2v_model):
        super().__init__()
        self.w2v_model = w2v_model

    def forward(self, source, padding_mask, **kwargs):
        """avhubert/hubert_asr.py:380-394: B x T x C -> T x B x C view + the three-key dict."""
        x, padding_mask = self.w2v_model.extract_finetune(source=source, padding_mask=padding_mask)
        x = x.transpose(0, 1)
        return {
            "encoder_out": x,                        # T x B x C
            "encoder_padding_mask": padding_mask,    # B x T
            "padding_mask": padding_mask,
        }

    def reorder_encoder_out(self, encoder_out, new_order):
        """avhubert/hubert_asr.py:396-409"""
        if encoder_out["encoder_out"] is not None:
            encoder_out["encoder_out"] = encoder_out["encoder_out"].index_select(1, new_order)
        if encoder_out["encoder_padding_mask"] is not None:
            encoder_out["encoder_padding_mask"] = encoder_out["encoder_padding_mask"].index_select(0, new_order)
        if encoder_out["padding_mask"] is not None:
            encoder_out["padding_mask"] = encoder_out["padding_mask"].index_select(0, new_order)
        return encoder_out
