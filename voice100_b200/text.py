"""Host-side tail of CTC greedy decoding (voice100/text.py:14,74-104): ids -> characters -> collapse."""
import re
from typing import Iterable

DEFAULT_CHARACTERS = "_ abcdefghijklmnopqrstuvwxyz'"
DEFAULT_VOCAB_SIZE = len(DEFAULT_CHARACTERS)


class CharTokenizer:
    def __init__(self, vocab=None):
        self._vocab = DEFAULT_CHARACTERS if vocab is None else vocab
        self.vocab_size = len(self._vocab)
        self._v2i = {ch: i for i, ch in enumerate(self._vocab)}

    def encode(self, text: str):
        import torch
        return torch.tensor([self._v2i[ch] for ch in text if ch in self._v2i], dtype=torch.long)

    def decode(self, encoded: Iterable[int]) -> str:
        return "".join(self._vocab[int(x)] for x in encoded if 0 <= int(x) < len(self._vocab))

    def merge_repeated(self, text: str) -> str:
        text = re.sub(r"(.)\1+", r"\1", text).replace("_", "")
        return "" if text == " " else text
