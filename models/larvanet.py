"""Lower-case alias so `--model=larvanet` (BASELINE.json north star) resolves on case-sensitive file systems."""
from models.LarvaNet import *  # noqa: F401,F403
from models.LarvaNet import create_model  # noqa: F401
