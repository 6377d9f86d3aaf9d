"""TEST INFRASTRUCTURE ONLY -- import stub: reference utils.py:3-4 imports nltk BLEU at module
import time; BLEU is outside the hot path."""
