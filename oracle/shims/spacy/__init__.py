"""TEST INFRASTRUCTURE ONLY -- import shim so that the UNMODIFIED /root/reference/data_processing.py imports in the build
container (spaCy is not installed and not installable offline).  data_processing.py:11 loads a pipeline at import time and only
the vocabulary builder / caption tokeniser (out of scope) ever calls it."""


class _Tok:
    def __init__(self, text):
        self.text = text


class _Pipeline:
    def tokenizer(self, text):
        return [_Tok(t) for t in str(text).split()]

    def __call__(self, text):
        return self.tokenizer(text)


def load(name, *a, **k):
    return _Pipeline()
