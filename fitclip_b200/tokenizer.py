"""``clip.tokenize(texts, truncate=True)`` hook (``aligner/encoder/clip_video_text_encoder.py:64-65``).

CLIP's BPE vocabulary (``bpe_simple_vocab_16e6.txt.gz``) ships with the un-vendored ``clip`` package and is not on
this image, so the tokenizer resolves lazily: the in-tree byte-pair tokenizer (``bpe.py``) on the merges file
``$FITCLIP_BPE_VOCAB`` points at, else the real ``clip`` package if importable, else a ``transformers.CLIPTokenizer``
built from ``$FITCLIP_CLIP_TOKENIZER`` (a directory with vocab.json / merges.txt).
The padding / truncation rule (SOT 49406, EOT 49407, zero pad to 77, overflow -> cut and force EOT last, int32) is
implemented here in :func:`pad_tokens`, and is what the tests pin."""
from __future__ import annotations

import os
from typing import Iterable, Iterator, List, Mapping, Sequence

import torch

SOT_TOKEN, EOT_TOKEN, CONTEXT_LENGTH = 49406, 49407, 77


def pad_tokens(token_lists: Sequence[Sequence[int]], context_length: int = CONTEXT_LENGTH,
               truncate: bool = True, sot: int = SOT_TOKEN, eot: int = EOT_TOKEN) -> torch.Tensor:
    """``[SOT] + ids + [EOT]`` rows -> int32 ``(n, context_length)``; twin ``aligner/encoder/slip.py:149-164``."""
    result = torch.zeros(len(token_lists), context_length, dtype=torch.int32)
    for i, ids in enumerate(token_lists):
        tokens = [sot, *ids, eot]
        if len(tokens) > context_length:
            if not truncate:
                raise RuntimeError(f"Input {i} is too long for context length {context_length}")
            tokens = tokens[:context_length]
            tokens[-1] = eot
        result[i, :len(tokens)] = torch.tensor(tokens, dtype=torch.int32)
    return result


_OWN = {}


def _bpe_encoder():
    vocab = os.environ.get("FITCLIP_BPE_VOCAB")  # bpe_simple_vocab_16e6.txt.gz: the in-tree tokenizer (bpe.py), no imports
    if vocab:
        if vocab not in _OWN:
            from .bpe import BpeTokenizer
            _OWN[vocab] = BpeTokenizer(vocab)
        return _OWN[vocab].encode, _OWN[vocab].decode
    try:
        from clip import clip as _clip  # noqa
        return lambda text: _clip._tokenizer.encode(text), lambda ids: _clip._tokenizer.decode(ids)
    except ImportError:
        pass
    path = os.environ.get("FITCLIP_CLIP_TOKENIZER")
    if path:
        from transformers import CLIPTokenizer
        tok = CLIPTokenizer.from_pretrained(path)
        return (lambda text: tok(text, add_special_tokens=False)["input_ids"]), (lambda ids: tok.decode(ids))
    raise RuntimeError("No CLIP BPE vocabulary available: point FITCLIP_BPE_VOCAB at bpe_simple_vocab_16e6.txt.gz, install "
                       "openai/CLIP, or point FITCLIP_CLIP_TOKENIZER at a directory holding the CLIP vocab.json/merges.txt "
                       "(synthetic benchmarks pass token ids directly)")


def tokenize(texts: Iterable[str]) -> Mapping[str, torch.Tensor]:
    """The encoder's ``get_tokenizer()`` return value. Module-level, hence picklable for DataLoader workers."""
    encode, _ = _bpe_encoder()
    if isinstance(texts, str):
        texts = [texts]
    own = _OWN.get(os.environ.get("FITCLIP_BPE_VOCAB") or "")
    if own is not None:  # the marker ids follow the merges file (49406 / 49407 with CLIP's)
        return {"input_ids": own.clip_tokenize(texts)}
    return {"input_ids": pad_tokens([encode(t) for t in texts])}


def slip_tokenize(texts: Iterable[str]) -> Mapping[str, torch.Tensor]:
    """``SlipVideoTextEncoder._tokenize`` (slip_video_text_encoder.py:53-54): the in-tree tokenizer's own framing -- int64,
    over-long rows cut without forcing EOT, a single text gives a 1-D row -- when ``$FITCLIP_BPE_VOCAB`` is set; otherwise
    the CLIP framing of :func:`tokenize` with whatever vocabulary source resolves."""
    vocab = os.environ.get("FITCLIP_BPE_VOCAB")
    if not vocab:
        return tokenize(texts)
    _bpe_encoder()
    return {"input_ids": _OWN[vocab].slip_tokenize(texts)}


def decode(input_ids: Iterable[Sequence[int]]) -> Iterator[str]:
    _, dec = _bpe_encoder()
    for ids in input_ids:
        yield dec([int(i) for i in ids])
