def sentence_bleu(*a, **k):
    raise NotImplementedError("nltk stub: BLEU is out of scope for the hot path")


class SmoothingFunction:
    method1 = None
