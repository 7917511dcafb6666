"""Caption auto-encoder pre-training: mirror of txt2vid/train/txt.py -- the sentence dataset and its
collate function (train/txt.py:21-52), one optimisation step (train/txt.py:166-181: encode -> teacher-forced or greedy
decode from the encoder's final state -> cross entropy against the padded targets) and the summed-loss evaluation
(train/txt.py:54-78).  Host logic plus the Seq2Seq modules of txt2vid_b200/text.py: embedding gather, engine GEMMs and the
length-masked recurrence kernels t2v_lstm_seq_fwd / _bwd, forward and backward (no cuDNN); parity against a fixture
recorded from the live reference: tests/test_round2_gpu.py::test_caption_pretraining_step_vs_reference_golden."""
import random

import torch
import torch.nn as nn
from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence


class SentenceDataset(torch.utils.data.Dataset):
    """train/txt.py:21-41: token index tensors of every caption (from a pickled {video: [captions]} dict or a list)."""

    def __init__(self, vocab=None, sent_path=None, sents=None):
        assert vocab is not None
        self.vocab, self.sent_path = vocab, sent_path
        if sent_path is not None:
            from .util import load
            temp = load(sent_path)
            self.sents = [s for x in temp for s in temp[x]]
        else:
            assert sents is not None
            self.sents = sents

    def __len__(self):
        return len(self.sents)

    def __getitem__(self, idx):
        return torch.Tensor([self.vocab(token) for token in self.vocab.tokenize(self.sents[idx])])


def collate_fn(data):
    """train/txt.py:44-52: longest first, zero padded LongTensor + lengths."""
    data.sort(key=lambda x: len(x), reverse=True)
    lengths = [len(sent) for sent in data]
    targets = torch.zeros(len(data), max(lengths)).long()
    for i, sent in enumerate(data):
        targets[i, :lengths[i]] = sent[:lengths[i]]
    return targets, lengths


def pretrain_step(seq2seq, sent, lengths, optimizer=None, teacher_force=True, criteria=None):
    """One iteration of train/txt.py:166-181.  Returns (loss, decoded symbols); steps `optimizer` when given."""
    criteria = criteria if criteria is not None else nn.CrossEntropyLoss()
    seq2seq.zero_grad()
    _, hidden_states, _ = seq2seq.encode(sent, lengths=lengths)
    packed = pack_padded_sequence(sent, lengths, batch_first=True)
    targets, _ = pad_packed_sequence(packed, batch_first=True, total_length=lengths[0])
    decoded, symbols = seq2seq.decode(true_inputs=sent, initial_hidden=hidden_states, max_seq_len=lengths[0],
                                      teacher_force=teacher_force)
    loss = criteria(decoded.permute(0, 2, 1), targets)
    loss.backward()
    if optimizer is not None:
        optimizer.step()
    return loss.detach(), symbols


def evaluate(split, seq2seq, device):
    """train/txt.py:54-78: teacher-free decoding, summed cross entropy per example."""
    seq2seq.eval()
    criteria = nn.CrossEntropyLoss(reduction='sum')
    loss, num_examples = 0.0, 0
    with torch.no_grad():
        for sent, lengths in split:
            sent = sent.to(device)
            _, hidden_states, _ = seq2seq.encode(sent, lengths=lengths)
            packed = pack_padded_sequence(sent, lengths, batch_first=True)
            targets, _ = pad_packed_sequence(packed, batch_first=True, total_length=lengths[0])
            decoded, _ = seq2seq.decode(true_inputs=sent, initial_hidden=hidden_states, max_seq_len=lengths[0],
                                        teacher_force=False)
            loss += float(criteria(decoded.permute(0, 2, 1), targets))
            num_examples += sent.size(0)
    seq2seq.train()
    return loss / max(1, num_examples)


def train(seq2seq, dataset, optimizer, device, epochs=1, teacher_force_prob=1.0, log=None):
    """The loop of train/txt.py:158-205 without its tensorboard / checkpoint side effects."""
    losses = []
    for _ in range(epochs):
        for sent, lengths in dataset:
            tf = random.uniform(0, 1) <= teacher_force_prob
            loss, symbols = pretrain_step(seq2seq, sent.to(device), lengths, optimizer, tf)
            losses.append(float(loss))
            if log is not None:
                log(len(losses), losses[-1], symbols)
    return losses
