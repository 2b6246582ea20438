"""Golden token ids from the REFERENCE'S OWN tokenizer (``aligner/encoder/slip.py:75-164``, ``SimpleTokenizer``), run in the
build container: ``python tests/golden/make_reference_bpe_golden.py`` -> ``tests/golden/bpe_synthetic_vocab.txt.gz`` +
``tests/golden/reference_bpe.pt``.

CLIP's merges file is not on disk, so the script first WRITES a synthetic one in the same format (a version line, then one
"left right" rule per line): rules learnt by a few hundred rounds of plain byte-pair counting over a small built-in corpus
(this trainer is the script's own; it only has to produce a plausible, deterministic rule list).  The reference class takes
the file's path as its constructor argument, so its real code runs on it unchanged."""
import collections
import gzip
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from make_reference_golden import REFERENCE, install_stubs  # noqa: E402

CORPUS = """
a man is playing a guitar on the street while people are watching him . a woman is cooking pasta in the kitchen and
talking to the camera ! two dogs are running across the field , chasing a red ball . the children's team won the game 3 - 2
in 2021 ; it's their first title . someone is slicing tomatoes , onions and peppers for a salad . a cartoon character
jumps over the wall and falls into the water . the news anchor is reporting about the weather : rain , snow and wind .
a person is folding a paper airplane and throwing it . people are dancing at a wedding party , they're laughing and singing .
a car is driving fast on the highway at night . the chef doesn't add salt , he'll add sugar instead . i've seen this
video 100 times , i'm sure you'd like it . naïve café déjà vu — señor , ¿ qué tal ? 日本語 のテキスト emoji 🙂 test
""".split()

SENTENCES = [
    "a man is playing a guitar", "A Woman is COOKING pasta!!", "it's the children's game, they're winning 3-2",
    "   lots   of\twhite\n\nspace   ", "don't, won't, I'll, you'd, we've, I'm", "naïve café — déjà vu, señor ¿qué tal?",
    "日本語のテキスト and emoji 🙂🙂", "&amp;lt;b&amp;gt; html &amp; entities &quot;quoted&quot;", "x", "", "1234567890 42",
    "<|startoftext|> markers inside <|endoftext|> text", "hyphen-ated under_scored dots...and???", "ⅻ ½ ٣ digits of other scripts",
    "zzzzqqqq unseenwordwithnomerges", " ".join(["a very long caption about people dancing"] * 20),
]


def learn_rules(words, rounds):
    """Plain BPE training on `words` (already mapped to the printable byte alphabet, last symbol carrying '</w>')."""
    vocab = collections.Counter(tuple(w) for w in words)
    rules = []
    for _ in range(rounds):
        pairs = collections.Counter()
        for word, n in vocab.items():
            for pair in zip(word, word[1:]):
                pairs[pair] += n
        if not pairs:
            break
        (a, b), _ = max(pairs.items(), key=lambda kv: (kv[1], kv[0]))
        rules.append((a, b))
        merged = collections.Counter()
        for word, n in vocab.items():
            out, i = [], 0
            while i < len(word):
                if i + 1 < len(word) and word[i] == a and word[i + 1] == b:
                    out.append(a + b)
                    i += 2
                else:
                    out.append(word[i])
                    i += 1
            merged[tuple(out)] += n
        vocab = merged
    return rules


def main() -> None:
    assert os.path.isdir(REFERENCE), f"{REFERENCE} is not mounted: this script only runs in the build container"
    install_stubs()
    sys.path.insert(0, REFERENCE)
    # slip.py evaluates default_bpe() at import (a path only, nothing is opened): import it as it is
    from aligner.encoder import slip as ref_slip  # noqa: E402

    table = ref_slip.bytes_to_unicode()
    words = []
    for w in CORPUS:
        chars = [table[b] for b in w.lower().encode("utf-8")]
        chars[-1] += "</w>"
        words.append(chars)
    rules = learn_rules(words, 400)
    vocab_path = os.path.join(HERE, "bpe_synthetic_vocab.txt.gz")
    with gzip.GzipFile(vocab_path, "wb", mtime=0) as f:  # mtime=0: byte-identical file on every run
        f.write(("#version: 0.2 (synthetic, tests only)\n" + "\n".join(f"{a} {b}" for a, b in rules) + "\n").encode("utf-8"))

    tok = ref_slip.SimpleTokenizer(bpe_path=vocab_path)
    out = {"sentences": SENTENCES, "reference_files": ["aligner/encoder/slip.py"], "num_rules": len(rules),
           "sot": tok.encoder["<|startoftext|>"], "eot": tok.encoder["<|endoftext|>"], "vocab_size": len(tok.encoder)}
    out["encoded"] = [tok.encode(s) for s in SENTENCES]
    out["decoded"] = [tok.decode(ids) for ids in out["encoded"]]
    out["batch_77"] = tok(SENTENCES)                       # (n, 77) int64, over-long rows cut (slip.py:145-164)
    out["batch_16"] = tok(SENTENCES, context_length=16)
    out["single"] = tok(SENTENCES[0])                      # 1-D
    out["byte_table"] = dict(table)
    path = os.path.join(HERE, "reference_bpe.pt")
    torch.save(out, path)
    print("wrote", vocab_path, os.path.getsize(vocab_path), "bytes;", path, os.path.getsize(path), "bytes;", len(rules), "rules")


if __name__ == "__main__":
    main()
