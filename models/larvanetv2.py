"""Lower-case alias so `--model=larvanetv2` resolves on case-sensitive file systems."""
from models.LarvaNetV2 import *  # noqa: F401,F403
from models.LarvaNetV2 import create_model  # noqa: F401
