"""Drop-in for the reference's `captioning.models` on the BoFi hot path:
`setup(opt)` (reference: captioning/models/__init__.py:14-24) returns a model whose
`state_dict()` layout, `forward(*args, mode=...)` dispatch and `_sample(fc, att, masks, opt)`
contract are those of the reference `TransformerModel` built with `train_mode: UIC`."""
from .TransformerModel import TransformerModel


def setup(opt):
    if getattr(opt, "caption_model", "transformer") == "transformer":
        return TransformerModel(opt)
    raise Exception("Caption model not supported: {}".format(opt.caption_model))
