"""Byte-pair tokenizer of CLIP, built from a merges file (``bpe_simple_vocab_16e6.txt.gz``) -- SURVEY.md 8 rows a8 / f4.

The reference tokenises with ``clip.tokenize`` (un-vendored ``clip`` package) and, for the SLIP-layout encoders, with its
in-tree ``SimpleTokenizer`` (``aligner/encoder/slip.py:75-164``); both read the same merges file, which is not on this
image.  This module restates the algorithm so that the encoders' ``get_tokenizer()`` hook works wherever that file is
(``FITCLIP_BPE_VOCAB=/path/to/bpe_simple_vocab_16e6.txt.gz``), with no ``clip`` / ``transformers`` import:

* vocabulary = the 256 byte symbols (GPT-2's printable re-mapping), the same 256 with the end-of-word mark ``</w>``, one
  entry per merge rule (the file's lines 2 .. 48895), then ``<|startoftext|>`` and ``<|endoftext|>`` -- ids follow that order
  (49406 / 49407 for the two markers with the real file);
* cleaning: ``ftfy.fix_text`` when ftfy is installed (it is not, here: skipped), ``html.unescape`` twice, whitespace runs
  collapsed, lower case;
* pre-tokenisation with CLIP's pattern (markers, English clitics, letter runs, single digits, other symbol runs), each
  piece mapped byte by byte to the printable alphabet, the last symbol carrying ``</w>``;
* merging: while some adjacent pair has a rank, take the lowest-ranked pair and fuse every non-overlapping occurrence of it,
  left to right.

Pinned against the reference's own ``SimpleTokenizer`` run on a synthetic merges file in the build container
(``tests/golden/make_reference_bpe_golden.py``, ``tests/test_bpe.py``).  Host code: tokenisation runs in DataLoader workers
in the reference too and is outside every timed region.
"""
from __future__ import annotations

import gzip
import html
from typing import Dict, Iterable, List, Sequence, Tuple, Union

import torch

END = "</w>"
SOT_TEXT, EOT_TEXT = "<|startoftext|>", "<|endoftext|>"
MAX_MERGES = 49152 - 256 - 2  # the reference keeps lines [1 : 48895) of the file (slip.py:80)
# CLIP's pre-tokenisation pattern (needs the `regex` package for \p classes), applied case-insensitively
PATTERN = r"""<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+"""


def byte_alphabet() -> Dict[int, str]:
    """byte value -> one printable character: the bytes that already print keep their own code point (``!``..``~``,
    ``¡``..``¬``, ``®``..``ÿ``), the other 68 are moved to 256, 257, ... in byte order (GPT-2's table, slip.py:28-48)."""
    printable = set(range(0x21, 0x7F)) | set(range(0xA1, 0xAD)) | set(range(0xAE, 0x100))
    ordered = sorted(printable, key=lambda b: (0 if b < 0x7F else 1 if b < 0xAD else 2, b))
    table = {b: chr(b) for b in ordered}
    spare = 256
    for b in range(256):
        if b not in printable:
            table[b] = chr(spare)
            spare += 1
    return table


def _clean(text: str) -> str:
    try:
        import ftfy  # the reference's basic_clean (slip.py:63-66) starts with ftfy.fix_text
        text = ftfy.fix_text(text)
    except ImportError:
        pass
    import regex
    text = html.unescape(html.unescape(text)).strip()
    return regex.sub(r"\s+", " ", text).strip().lower()  # slip.py:63-72,131


class BpeTokenizer:
    def __init__(self, bpe_path: str) -> None:
        import regex
        with gzip.open(bpe_path) as f:
            lines = f.read().decode("utf-8").split("\n")
        rules: List[Tuple[str, ...]] = [tuple(line.split()) for line in lines[1:1 + MAX_MERGES]]
        alphabet = byte_alphabet()
        # insertion order of `alphabet` is the vocabulary order of the single-byte symbols
        symbols = list(alphabet.values())
        vocab = symbols + [s + END for s in symbols] + ["".join(r) for r in rules] + [SOT_TEXT, EOT_TEXT]
        self.token_id: Dict[str, int] = {}
        for i, tok in enumerate(vocab):  # later duplicates win, as dict(zip(...)) makes them in the reference
            self.token_id[tok] = i
        self.id_token = {i: t for t, i in self.token_id.items()}
        self.rank = {}
        for i, r in enumerate(rules):
            self.rank[r] = i
        self.byte_char = alphabet
        self.char_byte = {c: b for b, c in alphabet.items()}
        self.splitter = regex.compile(PATTERN, regex.IGNORECASE)
        self.sot_token, self.eot_token = self.token_id[SOT_TEXT], self.token_id[EOT_TEXT]
        self._memo: Dict[str, List[str]] = {SOT_TEXT: [SOT_TEXT], EOT_TEXT: [EOT_TEXT]}

    def _merge(self, piece: str) -> List[str]:
        """Symbols of one pre-token after all applicable merges."""
        hit = self._memo.get(piece)
        if hit is not None:
            return hit
        word = list(piece[:-1]) + [piece[-1] + END]
        while len(word) > 1:
            best, best_rank = None, None
            for pair in zip(word, word[1:]):
                r = self.rank.get(pair)
                if r is not None and (best_rank is None or r < best_rank):
                    best, best_rank = pair, r
            if best is None:
                break
            fused, out, i = best[0] + best[1], [], 0
            while i < len(word):
                if i + 1 < len(word) and word[i] == best[0] and word[i + 1] == best[1]:
                    out.append(fused)
                    i += 2
                else:
                    out.append(word[i])
                    i += 1
            word = out
        self._memo[piece] = word
        return word

    def encode(self, text: str) -> List[int]:
        ids: List[int] = []
        for piece in self.splitter.findall(_clean(text)):
            mapped = "".join(self.byte_char[b] for b in piece.encode("utf-8"))
            ids.extend(self.token_id[s] for s in self._merge(mapped))
        return ids

    def decode(self, tokens: Iterable[int]) -> str:
        # every character back to its byte, then the end-of-word marks become spaces (slip.py:140-143; "</w>" is printable
        # ASCII, which the byte table maps to itself)
        chars = "".join(self.id_token[int(t)] for t in tokens)
        return bytearray(self.char_byte[c] for c in chars).decode("utf-8", errors="replace").replace(END, " ")

    # ---- the two framing rules of the reference ---------------------------------------------------------------------
    def clip_tokenize(self, texts: Union[str, Sequence[str]], context_length: int = 77, truncate: bool = True) -> torch.Tensor:
        """``clip.tokenize`` [3P]: int32 ``(n, context_length)``; an over-long row is cut and its last id forced to EOT."""
        from .tokenizer import pad_tokens
        texts = [texts] if isinstance(texts, str) else list(texts)
        return pad_tokens([self.encode(t) for t in texts], context_length, truncate, sot=self.sot_token, eot=self.eot_token)

    def slip_tokenize(self, texts: Union[str, Sequence[str]], context_length: int = 77) -> torch.Tensor:
        """``SimpleTokenizer.__call__`` (slip.py:145-164): int64, an over-long row is simply cut (EOT is lost), and a single
        text gives a 1-D tensor."""
        texts = [texts] if isinstance(texts, str) else list(texts)
        out = torch.zeros(len(texts), context_length, dtype=torch.long)
        for i, t in enumerate(texts):
            ids = ([self.sot_token] + self.encode(t) + [self.eot_token])[:context_length]
            out[i, :len(ids)] = torch.tensor(ids)
        return out[0] if len(out) == 1 else out

    __call__ = slip_tokenize
