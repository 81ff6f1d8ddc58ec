"""Model hyper-parameters and the reference checkpoint layout for the BoFi hot path.

`BofiConfig.from_opt` reads exactly the option names the reference model reads
(/root/reference/captioning/models/AttModel.py:56-104, TransformerModel.py:1631-1640, :404-411),
with the same getattr defaults, so an `infos.pkl['opt']` namespace or a YAML-derived namespace
configures this implementation unchanged.

`state_spec` enumerates the reference's `state_dict()` in the reference's
own order: (name, shape, kind).  It is the single source of truth for the drop-in module tree,
the synthetic checkpoint generator and the device weight packer.
"""
from collections import OrderedDict
from dataclasses import dataclass, asdict

LENGTH_DIM = 20      # TransformerModel.py:329
SYN_DIM = 10         # TransformerModel.py:330
SYN_LOWER = 4        # TransformerModel.py:331 (VP=4, NP=5, CP=6)
SYN_UPPER = 6        # TransformerModel.py:332
HEAD_HIDDEN = 100    # TransformerModel.py:346-349
PE_MAX_LEN = 5000    # TransformerModel.py:1491


@dataclass
class BofiConfig:
    vocab_size: int = 9487
    att_feat_size: int = 2048
    N_enc: int = 6
    N_dec: int = 6
    N_len: int = 1
    d_model: int = 512
    d_ff: int = 2048
    h: int = 8
    seq_length: int = 20
    max_boxes: int = 100
    dropout: float = 0.1
    drop_prob_lm: float = 0.5
    pad_idx: int = 0
    bos_idx: int = 1
    eos_idx: int = 2
    len_idx: int = 3
    decoder_input_mode: str = "add"
    train_mode: str = "UIC"

    @property
    def tgt_vocab(self):
        return self.vocab_size + 4

    @property
    def bound_slots(self):
        return self.seq_length + 2

    def to_dict(self):
        return asdict(self)

    @classmethod
    def from_opt(cls, opt):
        g = lambda k, d=None: getattr(opt, k, d)
        if g("use_bn", 0):
            raise NotImplementedError("use_bn != 0 is outside the uic_sd hot path")
        if g("caption_model", "transformer") != "transformer":
            raise Exception("Caption model not supported: {}".format(g("caption_model")))
        train_mode = g("train_mode", "AIC")
        if train_mode != "UIC":
            raise NotImplementedError("only train_mode=UIC (the BoFi model of uic_sd*.yml) is built; got %r" % train_mode)
        mode = g("decoder_input_mode", "add")
        if mode != "add":
            raise NotImplementedError("decoder_input_mode=%r (uic_sd*.yml use 'add')" % mode)
        return cls(
            vocab_size=int(opt.vocab_size),
            att_feat_size=int(g("att_feat_size", 2048)),
            N_enc=int(g("N_enc", g("num_layers", 6))),
            N_dec=int(g("N_dec", g("num_layers", 6))),
            N_len=int(g("N_len", 0)),
            d_model=int(g("d_model", g("input_encoding_size", 512))),
            d_ff=int(g("d_ff", g("rnn_size", 2048))),
            h=int(g("num_att_heads", 8)),
            seq_length=int(g("max_length", 20) or g("seq_length", 20)),
            max_boxes=int(g("max_boxes", 100)),
            dropout=float(g("dropout", 0.1)),
            drop_prob_lm=float(g("drop_prob_lm", 0.5)),
            pad_idx=int(g("pad_idx", 0)), bos_idx=int(g("bos_idx", 1)),
            eos_idx=int(g("eos_idx", 2)), len_idx=int(g("len_idx", 3)),
            decoder_input_mode=mode, train_mode=train_mode)


def _attn(prefix, d):
    for i in range(4):
        yield prefix + ".linears.%d.weight" % i, (d, d), "matrix"
        yield prefix + ".linears.%d.bias" % i, (d,), "bias"


def _ffn(prefix, d, dff):
    yield prefix + ".w_1.weight", (dff, d), "matrix"
    yield prefix + ".w_1.bias", (dff,), "bias"
    yield prefix + ".w_2.weight", (d, dff), "matrix"
    yield prefix + ".w_2.bias", (d,), "bias"


def _norm(prefix, d):
    yield prefix + ".a_2", (d,), "ones"
    yield prefix + ".b_2", (d,), "zeros"


def _stack_layer(prefix, d, dff, cross, ffname):
    yield from _attn(prefix + ".self_attn", d)
    if cross:
        yield from _attn(prefix + ".src_attn", d)
    yield from _ffn(prefix + "." + ffname, d, dff)
    for s in range(3 if cross else 2):
        yield from _norm(prefix + ".sublayer.%d.norm" % s, d)


def _linear(prefix, out_f, in_f):
    yield prefix + ".weight", (out_f, in_f), "matrix"
    yield prefix + ".bias", (out_f,), "bias"


def state_spec(cfg):
    """OrderedDict name -> (shape, kind) ; kind in matrix|bias|ones|zeros|embedding|pe."""
    d, dff, V = cfg.d_model, cfg.d_ff, cfg.tgt_vocab
    items = []
    items += _linear("att_embed.0", d, cfg.att_feat_size)
    for l in range(cfg.N_enc):
        items += _stack_layer("model.encoder.layers.%d" % l, d, dff, False, "feed_forward")
    items += _norm("model.encoder.norm", d)
    for l in range(cfg.N_dec):
        items += _stack_layer("model.decoder.layers.%d" % l, d, dff, True, "feed_forward")
    items += _norm("model.decoder.norm", d)
    items.append(("model.syn_embed.lut.weight", (SYN_DIM, d), "embedding"))
    items.append(("model.tgt_embed.lut.weight", (V, d), "embedding"))
    items.append(("model.pos_embed.pe", (1, PE_MAX_LEN, d), "pe"))
    items += _linear("model.generator.proj", V, d)
    lp = "model.length_predictor"
    items += _attn(lp + ".length_attn", d)
    items += _ffn(lp + ".ff", d, dff)
    items += _norm(lp + ".norm", d)
    items += _linear(lp + ".Length_classifier1", HEAD_HIDDEN, d)
    items += _linear(lp + ".Length_classifier2", LENGTH_DIM, HEAD_HIDDEN)
    items += _linear(lp + ".Syntactic_classifier1", HEAD_HIDDEN, d)
    items += _linear(lp + ".Syntactic_classifier2", SYN_DIM, HEAD_HIDDEN)
    if cfg.N_len == 0:
        items += _norm(lp + ".LengthPredictor.norm", d)
    else:
        for l in range(cfg.N_len):
            items += _stack_layer(lp + ".LengthPredictor.%d" % l, d, dff, True, "ff")
    return OrderedDict((n, (tuple(s), k)) for n, s, k in items)
