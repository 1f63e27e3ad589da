"""Host-side tail of CTC greedy decoding: ids -> symbols -> collapse (voice100/text.py:14-44,74-145).

`CharTokenizer` serves the character vocabularies (asr_en_*: V = 29), `BasicTokenizer` the phone vocabularies
(asr_en_phone_*: 71 ARPAbet symbols joined by '/', asr_ja_phone_*: 44 symbols joined by ' ').  Only
encode / decode / merge_repeated are on the inference path; the phonemizers (g2p_en, pyopenjtalk) are front ends
outside it.  `merge_repeated` works on the token list (run-length collapse, then drop the blank) -- the same
result as the reference's regular expressions for every string `decode` can produce (tests/golden/tokenizer.json).
"""
from itertools import groupby
from typing import Iterable, List, Optional

DEFAULT_CHARACTERS = "_ abcdefghijklmnopqrstuvwxyz'"
DEFAULT_VOCAB_SIZE = len(DEFAULT_CHARACTERS)

# ARPAbet in alphabetical order; the 15 vowels carry a stress digit (CMUdict), consonants do not; the inventory the
# reference trains on additionally has a bare 'UW' (voice100/text.py:19-30).  Index 0 is the CTC blank.
_ARPABET = ("AA AE AH AO AW AY B CH D DH EH ER EY F G HH IH IY JH K L M N NG OW OY P R S SH T TH UH UW V W Y Z ZH").split()
_ARPABET_VOWELS = set("AA AE AH AO AW AY EH ER EY IH IY OW OY UH UW".split())


def _cmu_vocab() -> List[str]:
    out = ["_"]
    for ph in _ARPABET:
        if ph in _ARPABET_VOWELS:
            out += ([ph] if ph == "UW" else []) + [ph + d for d in "012"]
        else:
            out.append(ph)
    return out


# Japanese phone set of pyopenjtalk g2p (blank '-', then punctuation, the moraic nasal 'N' and the phones in ASCII
# order, voice100/text.py:34-40)
_JA_PHONES = ("a a: b by ch d e e: f g gy h hy i i: j k ky m my n ny o o: p py q r ry s sh t ts u u: w y z").split()

CMU_VOCAB = _cmu_vocab()
JA_VOCAB = ["-"] + sorted(["!", ",", ".", "?", "N"] + _JA_PHONES)
assert len(CMU_VOCAB) == 71 and len(JA_VOCAB) == 44


def _ids(encoded: Iterable) -> List[int]:
    return [int(x) for x in encoded]


class CharTokenizer:
    """One character = one token (voice100/text.py:74-104)."""

    def __init__(self, vocab: Optional[str] = None):
        self._vocab = DEFAULT_CHARACTERS if vocab is None else vocab
        self.vocab_size = len(self._vocab)
        self._v2i = {ch: i for i, ch in enumerate(self._vocab)}

    def __call__(self, text: str):
        return self.encode(text)

    def encode(self, text: str):
        import torch
        return torch.tensor([self._v2i[ch] for ch in text if ch in self._v2i], dtype=torch.long)

    def decode(self, encoded: Iterable[int]) -> str:
        return "".join(self._vocab[x] for x in _ids(encoded) if 0 <= x < len(self._vocab))

    def merge_repeated(self, text: str) -> str:
        # runs of one character collapse to that character ('.' in the reference's regex does not match a newline)
        text = "".join(ch * len(list(g)) if ch == "\n" else ch for ch, g in groupby(text))
        text = text.replace("_", "")
        return "" if text == " " else text


class BasicTokenizer:
    """Separator-joined phone symbols (voice100/text.py:107-145): language 'en' = ARPAbet joined by '/',
    'ja' = Japanese phones joined by ' '."""

    def __init__(self, language: str):
        if language == "en":
            vocab, separator = CMU_VOCAB, "/"
        elif language == "ja":
            vocab, separator = JA_VOCAB, " "
        else:
            raise ValueError(f"unknown language {language!r} (expected 'en' or 'ja')")
        self.vocab_size = len(vocab)
        self._separator, self._vocab = separator, vocab
        self._v2i = {x: i for i, x in enumerate(vocab)}

    def __call__(self, text: str):
        return self.encode(text)

    def encode(self, text: str):
        import torch
        return torch.tensor([self._v2i[ph] for ph in text.split(self._separator) if ph in self._v2i], dtype=torch.long)

    def decode(self, encoded: Iterable[int]) -> str:
        return self._separator.join(self._vocab[x] for x in _ids(encoded) if 0 <= x < len(self._vocab))

    def merge_repeated(self, text: str) -> str:
        """Collapse runs of the same symbol, then drop the blank (the CTC rule on symbol strings)."""
        blank = self._vocab[0]
        runs = [ph for ph, _ in groupby(text.split(self._separator))]
        return self._separator.join(ph for ph in runs if ph != blank and ph != "")

    def decode_collapsed(self, ids: Iterable[int]) -> str:
        """ids already collapsed on the device (v100_ctc_collapse / AsrPipeline.transcribe_ids) -> text."""
        return self.decode(ids)
