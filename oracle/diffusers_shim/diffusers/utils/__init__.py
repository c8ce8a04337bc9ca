import logging as _pylogging

USE_PEFT_BACKEND = False
WEIGHTS_NAME = "diffusion_pytorch_model.bin"


class _Logging:
    @staticmethod
    def get_logger(name):
        return _pylogging.getLogger(name)


logging = _Logging()


def is_torch_version(op, ver):
    return True


def scale_lora_layers(model, weight):
    return None


def unscale_lora_layers(model, weight=None):
    return None


def load_image(x, *a, **k):  # pragma: no cover - image IO is outside the hot path
    from PIL import Image

    return Image.open(x) if isinstance(x, str) else x


def replace_example_docstring(doc):
    def deco(fn):
        return fn

    return deco
