"""Caption encoder (Embedding + 4-layer Bi-LSTM), mirror of txt2vid/models/txt/basic.py.

Same constructor kwargs, sub-module names (encoder.embed / encoder.lstm / encoder.to_vocab; decoder is
the same object when separate_decoder=False) and return structure of `encode`.  The nn.Embedding / nn.LSTM /
nn.Linear members are PARAMETER CONTAINERS (identical state_dict keys and initial weights); their forward is never
called on the GPU path:

  * token gather: t2v_embedding_fwd (bit-exact row copy; gradient by atomic row adds);
  * per layer, the input projection of all time steps and both directions is ONE GEMM on the tcgen05 engine, the
    sequential part is the persistent length-masked recurrence kernel t2v_lstm_seq_fwd (csrc/lstm.cu), which
    reproduces cuDNN's PackedSequence semantics (models/txt/basic.py:52-56): a sample advances only while t < len,
    the reverse direction starts at the sample's own last token, padded steps emit zeros;
  * backward (end-to-end training, caption pre-training): t2v_lstm_seq_bwd + three GEMMs + one column sum.
"""
import torch
import torch.nn as nn

from . import ops


def _lstm_params(lstm, layer, ndir):
    out = []
    for d in range(ndir):
        sfx = "_l%d%s" % (layer, "_reverse" if d == 1 else "")
        out += [getattr(lstm, "weight_ih" + sfx), getattr(lstm, "weight_hh" + sfx), getattr(lstm, "bias_ih" + sfx),
                getattr(lstm, "bias_hh" + sfx)]
    return out


def lstm_forward(lstm, x, lengths_dev, initial_state=None):
    """nn.LSTM(batch_first=True) container applied to x (B, L, In) storage-dtype tensor with per-sample lengths
    (int32 device tensor) -> (out (B, L, ndir*H) storage dtype, (h_n, c_n) fp32 (layers*ndir, B, H))."""
    ndir = 2 if lstm.bidirectional else 1
    hs, cs = [], []
    inp = x
    for layer in range(lstm.num_layers):
        h0 = c0 = None
        if initial_state is not None:
            h0 = initial_state[0][layer * ndir:(layer + 1) * ndir]
            c0 = initial_state[1][layer * ndir:(layer + 1) * ndir]
        inp, hn, cn = ops.LstmLayerF.apply(inp.contiguous(), lengths_dev, h0, c0, *_lstm_params(lstm, layer, ndir))
        hs.append(hn)
        cs.append(cn)
    return inp, (torch.cat(hs, dim=0), torch.cat(cs, dim=0))


class RecurrentModel(nn.Module):
    def __init__(self, vocab_size=None, embed_size=256, hidden_size=256, encoding_size=256, num_layers=4, bi=True,
                 is_decoder=False):
        super().__init__()
        self.bi = bi
        self.num_layers = num_layers
        self.hidden_size = hidden_size // 2 if bi else hidden_size
        self.encoding_size = encoding_size
        self.vocab_size = vocab_size
        self.embed_size = embed_size
        self.embed = nn.Embedding(vocab_size, embed_size)
        self.lstm = nn.LSTM(embed_size, self.hidden_size, num_layers, batch_first=True, bidirectional=bi)
        self.is_decoder = is_decoder
        if is_decoder:
            self.to_vocab = nn.Linear(hidden_size, vocab_size)

    def _to_vocab(self, h):
        """nn.Linear(hidden, vocab) on a (B, L, hidden) storage-dtype tensor -> fp32 (B, L, V), on the conv engine"""
        B, L, Hh = h.shape
        V = self.to_vocab.weight.shape[0]
        y = ops.conv(h.reshape(B * L, 1, 1, 1, Hh), self.to_vocab.weight, self.to_vocab.bias)
        return ops.SliceF32F.apply(y, V).reshape(B, L, V)

    def forward(self, x, lengths=None, initial_state=None, raw_output=True):
        """tokens (B,L) int64 sorted by length (desc) -> (out (B,L0,2H), (h,c), hn (B,2H))
        (txt/basic.py:49-70): hn = cat(final h of the last layer, forward and backward)."""
        B = x.shape[0]
        L0 = int(lengths[0])
        lens = torch.as_tensor([int(v) for v in lengths], dtype=torch.int32, device=x.device)
        emb = ops.EmbeddingF.apply(x[:, :L0].contiguous(), self.embed.weight)          # (B, L0, E) storage dtype
        seq, hidden = lstm_forward(self.lstm, emb, lens, initial_state)
        ndir = 2 if self.bi else 1
        last = hidden[0].view(self.num_layers, ndir, -1, self.hidden_size)[-1]
        hn = torch.cat((last[0], last[1]), dim=1) if self.bi else last[0]
        if not raw_output:
            assert self.is_decoder
            out = self._to_vocab(seq)
            if out.shape[1] == 1:
                out = out.squeeze(1)
        else:
            out = ops.SliceF32F.apply(seq, seq.shape[-1])                               # fp32 view for the caller
        return out, hidden, hn

    def sample(self, true_inputs=None, initial_hidden=None, max_seq_len=60, teacher_force=False):
        """Greedy / teacher-forced decoding (txt/basic.py:73-101); caption pre-training only."""
        logits, symbols = [], []
        inputs = true_inputs[:, 0].unsqueeze(1)
        hidden = initial_hidden
        B = true_inputs.shape[0]
        ones = torch.ones(B, dtype=torch.int32, device=true_inputs.device)
        for i in range(max_seq_len):
            emb = ops.EmbeddingF.apply(inputs.contiguous(), self.embed.weight)
            seq, hidden = lstm_forward(self.lstm, emb, ones, hidden)
            o = self._to_vocab(seq).squeeze(1)
            pred = o.max(1)[1]
            logits.append(o)
            symbols.append(pred)
            inputs = true_inputs[:, i].unsqueeze(1) if teacher_force else pred.unsqueeze(1)
        return torch.stack(logits, 1), torch.stack(symbols, 1)

    def create_initial_state(self):
        return torch.zeros(self.num_layers, 1, self.hidden_size)


class Seq2Seq(nn.Module):
    def __init__(self, separate_decoder=False, vocab_size=None):
        super().__init__()
        self.separate_decoder = separate_decoder
        self.encoder = RecurrentModel(vocab_size=vocab_size, is_decoder=not separate_decoder)
        self.decoder = RecurrentModel(vocab_size=vocab_size, is_decoder=True, bi=False) if separate_decoder \
            else self.encoder

    def encode(self, *args, **kwargs):
        return self.encoder(*args, **kwargs)

    def decode(self, *args, **kwargs):
        return self.decoder.sample(*args, **kwargs)
