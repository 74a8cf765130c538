"""Model factory with the reference's configuration contract (utils/get_model.py:4-12)."""
import importlib

_MODELS_PACKAGE = __name__.rsplit(".utils.", 1)[0] + ".models"


def get_model(config, device="cuda"):
    """Build the model a YAML ``model:`` block describes.

    ``config["script"]`` is the module name under ``models`` and ``config["class_name"]`` the class defined there (the two
    keys the reference's configs carry); the instance is constructed from the whole block and moved to ``device``."""
    try:
        script, class_name = config["script"], config["class_name"]
    except KeyError as missing:
        raise KeyError(f"model config needs 'script' and 'class_name' (missing {missing})") from None
    model_cls = getattr(importlib.import_module(f"{_MODELS_PACKAGE}.{script}"), class_name)
    return model_cls(config).to(device)
