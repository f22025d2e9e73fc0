"""CLIP byte-level BPE tokenizer (API surface of reference jclip/simple_tokenizer.py).

Text is outside the hot path; this exists so `clip.tokenize` keeps working when a
merges file is available.  The vocabulary file is NOT shipped with this repo: point
`JCLIP_BPE_VOCAB` (or the `bpe_path` argument) at OpenAI's `bpe_simple_vocab_16e6.txt[.gz]`.
The reference calls `ftfy.fix_text` first; when ftfy is not installed that step is skipped.
"""
import gzip
import html
import os
from functools import lru_cache

import regex as re

try:
    import ftfy
except ImportError:   # not installable offline
    ftfy = None

SOT, EOT = "<|startoftext|>", "<|endoftext|>"
N_MERGES = 49152 - 256 - 2


@lru_cache()
def byte_alphabet():
    """Printable stand-ins for the 256 byte values (GPT-2 / CLIP convention)."""
    keep = [*range(33, 127), *range(161, 173), *range(174, 256)]
    table, extra = {}, 0
    for b in keep:
        table[b] = chr(b)
    for b in range(256):
        if b not in table:
            table[b] = chr(256 + extra)
            extra += 1
    return table


def default_bpe():
    return os.environ.get("JCLIP_BPE_VOCAB", os.path.join(os.path.dirname(os.path.abspath(__file__)),
                                                          "bpe_simple_vocab_16e6.txt.gz"))


def _read_merges(path):
    with open(path, "rb") as f:
        raw = f.read()
    if raw[:2] == b"\x1f\x8b":
        raw = gzip.decompress(raw)
    lines = raw.decode("utf-8").split("\n")
    return [tuple(l.split()) for l in lines[1:N_MERGES + 1]]


class SimpleTokenizer:
    def __init__(self, bpe_path=None):
        path = bpe_path or default_bpe()
        if not os.path.exists(path):
            raise FileNotFoundError(f"BPE vocabulary {path} not found; set JCLIP_BPE_VOCAB")
        merges = _read_merges(path)
        alphabet = byte_alphabet()
        # insertion order of byte_alphabet() is keep-list first, then the remapped bytes: the
        # published vocabulary order
        symbols = list(alphabet.values())
        vocab = symbols + [s + "</w>" for s in symbols] + ["".join(m) for m in merges] + [SOT, EOT]
        self.encoder = {tok: i for i, tok in enumerate(vocab)}
        self.decoder = {i: tok for tok, i in self.encoder.items()}
        self.byte_encoder = alphabet
        self.byte_decoder = {c: b for b, c in alphabet.items()}
        self.rank = {m: i for i, m in enumerate(merges)}
        self.cache = {SOT: SOT, EOT: EOT}
        self.pat = re.compile(r"<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+",
                              re.IGNORECASE)

    def bpe(self, token):
        hit = self.cache.get(token)
        if hit is not None:
            return hit
        word = list(token[:-1]) + [token[-1] + "</w>"]
        while len(word) > 1:
            best, best_rank = None, None
            for pair in zip(word, word[1:]):
                r = self.rank.get(pair)
                if r is not None and (best_rank is None or r < best_rank):
                    best, best_rank = pair, r
            if best is None:
                break
            merged, i = [], 0
            while i < len(word):
                if i + 1 < len(word) and word[i] == best[0] and word[i + 1] == best[1]:
                    merged.append(best[0] + best[1])
                    i += 2
                else:
                    merged.append(word[i])
                    i += 1
            word = merged
        out = " ".join(word)
        self.cache[token] = out
        return out

    def encode(self, text):
        if ftfy is not None:
            text = ftfy.fix_text(text)
        text = html.unescape(html.unescape(text)).strip()
        text = re.sub(r"\s+", " ", text).strip().lower()
        ids = []
        for tok in re.findall(self.pat, text):
            tok = "".join(self.byte_encoder[b] for b in tok.encode("utf-8"))
            ids.extend(self.encoder[t] for t in self.bpe(tok).split(" "))
        return ids

    def decode(self, tokens):
        text = "".join(self.decoder[int(t)] for t in tokens)
        return bytearray(self.byte_decoder[c] for c in text).decode("utf-8", errors="replace").replace("</w>", " ")
