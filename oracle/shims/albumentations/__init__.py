"""TEST INFRASTRUCTURE ONLY -- import shim: /root/reference/dataset.py does `import albumentations as A` at import time and uses it
only inside the (out-of-scope) data-loading transforms."""


class _Missing:
    def __init__(self, *a, **k):
        raise RuntimeError("albumentations is not installed; the data-loading transforms are out of scope")


Compose = Resize = Normalize = BboxParams = HorizontalFlip = RandomBrightnessContrast = _Missing


def __getattr__(name):
    return _Missing
