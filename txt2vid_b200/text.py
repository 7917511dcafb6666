"""Caption encoder (Embedding + 4-layer Bi-LSTM), mirror of txt2vid/models/txt/basic.py.

Same constructor kwargs, sub-module names (encoder.embed / encoder.lstm / encoder.to_vocab; decoder is
the same object when separate_decoder=False) and return structure of `encode`.

ROUND-1 INTERIM: the recurrence itself still runs through the nn.LSTM container (cuDNN) -- 0.18 of the
443 GFLOP of an iteration and, with --end2end off (scripts/run_tganv2_cond.sh), forward-only.  The fused
persistent LSTM kernel (t2v_lstm_cell_* + the gate GEMM on the conv engine) is wired for the ConvLSTM
first; DESIGN.md tracks the Bi-LSTM port.
"""
import torch
import torch.nn as nn
from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence


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

    def forward(self, x, lengths=None, initial_state=None, raw_output=True):
        """tokens (B,L) int64 sorted by length (desc) -> (out (B,L0,2H), (h,c), hn (B,2H))
        (txt/basic.py:49-70): hn = cat(final h of the last layer, forward and backward)."""
        packed = pack_padded_sequence(self.embed(x), lengths, batch_first=True)
        out, hidden = self.lstm(packed, initial_state)
        out, _ = pad_packed_sequence(out, batch_first=True, total_length=lengths[0])
        ndir = 2 if self.bi else 1
        last = hidden[0].view(self.num_layers, ndir, -1, self.hidden_size)[-1]
        hn = torch.cat((last[0], last[1]), dim=1) if self.bi else last[0]
        if not raw_output:
            assert self.is_decoder
            out = self.to_vocab(out.squeeze(1))
        return out, hidden, hn

    def sample(self, true_inputs=None, initial_hidden=None, max_seq_len=60, teacher_force=False):
        """Greedy / teacher-forced decoding (txt/basic.py:73-101); caption pre-training only."""
        logits, symbols = [], []
        inputs = true_inputs[:, 0].unsqueeze(1)
        hidden = initial_hidden
        for i in range(max_seq_len):
            o, hidden = self.lstm(self.embed(inputs), hidden)
            o = self.to_vocab(o.squeeze(1))
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
