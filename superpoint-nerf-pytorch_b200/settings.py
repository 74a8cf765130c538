"""Path constants of the reference's generated ``superpoint/settings.py`` (superpoint/setup.sh:1-8), taken from the
environment instead of an interactive prompt."""
import os

DATA_PATH = os.environ.get("SPN_DATA_PATH", os.path.expanduser("~/spn_data"))
CKPT_PATH = os.environ.get("SPN_CKPT_PATH", os.path.expanduser("~/spn_ckpt"))
EXPER_PATH = os.environ.get("SPN_EXPER_PATH", os.path.expanduser("~/spn_exper"))
