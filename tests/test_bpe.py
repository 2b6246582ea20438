"""The in-tree byte-pair tokenizer (fitclip_b200/bpe.py) against the reference's own ``SimpleTokenizer``
(``aligner/encoder/slip.py:75-164``) run on a synthetic merges file in the build container
(tests/golden/make_reference_bpe_golden.py): ids, decoding, both framing rules.  Bit-exact (integer work)."""
import os

import pytest
import torch

from fitclip_b200 import tokenizer
from fitclip_b200.bpe import BpeTokenizer, byte_alphabet

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
VOCAB = os.path.join(GOLDEN, "bpe_synthetic_vocab.txt.gz")


@pytest.fixture(scope="module")
def ref():
    return torch.load(os.path.join(GOLDEN, "reference_bpe.pt"), map_location="cpu", weights_only=False)


@pytest.fixture(scope="module")
def tok():
    return BpeTokenizer(VOCAB)


def test_byte_alphabet_and_vocabulary_layout(ref, tok):
    assert byte_alphabet() == ref["byte_table"]
    assert list(byte_alphabet().values()) == list(ref["byte_table"].values())  # the ORDER is the id order
    assert (tok.sot_token, tok.eot_token) == (ref["sot"], ref["eot"])
    assert len(tok.token_id) == ref["vocab_size"]


def test_encode_matches_reference_on_every_sentence(ref, tok):
    for text, expect in zip(ref["sentences"], ref["encoded"]):
        assert tok.encode(text) == expect, text
    assert any(len(e) == 0 for e in ref["encoded"]) and max(len(e) for e in ref["encoded"]) > 77  # empty and over-long cases


def test_decode_matches_reference(ref, tok):
    for ids, expect in zip(ref["encoded"], ref["decoded"]):
        assert tok.decode(ids) == expect


def test_slip_framing_matches_reference_call(ref, tok):
    got = tok(ref["sentences"])
    assert got.dtype == torch.long and torch.equal(got, ref["batch_77"])
    assert torch.equal(tok(ref["sentences"], context_length=16), ref["batch_16"])
    single = tok(ref["sentences"][0])
    assert single.dim() == 1 and torch.equal(single, ref["single"])
    long_row = ref["batch_77"][-1]
    assert long_row[-1] != ref["eot"]  # SimpleTokenizer cuts over-long rows without forcing EOT (slip.py:158-160)


def test_clip_framing_forces_eot_and_int32(ref, tok):
    got = tok.clip_tokenize(ref["sentences"])
    assert got.dtype == torch.int32 and got.shape == (len(ref["sentences"]), 77)
    assert got[-1, -1] == ref["eot"] and got[-1, 0] == ref["sot"]  # [3P] clip.tokenize(truncate=True)
    short = ref["encoded"][0]
    assert got[0, :len(short) + 2].tolist() == [ref["sot"], *short, ref["eot"]] and got[0, len(short) + 2:].sum() == 0
    with pytest.raises(RuntimeError):
        tok.clip_tokenize(ref["sentences"][-1:], truncate=False)


def test_tokenizer_hooks_use_the_vocabulary_file(ref, monkeypatch):
    monkeypatch.setenv("FITCLIP_BPE_VOCAB", VOCAB)
    out = tokenizer.slip_tokenize(ref["sentences"])
    assert torch.equal(out["input_ids"], ref["batch_77"])
    assert list(tokenizer.decode([ref["encoded"][2]])) == [ref["decoded"][2]]
    clip_rule = tokenizer.tokenize(ref["sentences"])["input_ids"]
    assert clip_rule.dtype == torch.int32 and clip_rule[-1, -1] == ref["eot"] and clip_rule[0, 0] == ref["sot"]
    monkeypatch.delenv("FITCLIP_BPE_VOCAB")
    with pytest.raises(RuntimeError, match="FITCLIP_BPE_VOCAB"):
        tokenizer.tokenize(["no vocabulary anywhere"])
