import importlib


def get_model(config, device="cuda"):
    """Same contract as the reference's utils/get_model.py:4-12: ``config['script']`` names a module under
    ``models`` and ``config['class_name']`` the class inside it; returns ``Class(config).to(device)``."""
    pkg = __name__.rsplit(".utils.", 1)[0]
    module = importlib.import_module(f"{pkg}.models.{config['script']}")
    return getattr(module, config["class_name"])(config).to(device)
