"""Model-plugin contract (mirrors reference models/base.py:1-84).

Entry scripts discover a model with `importlib.import_module('models.' + name).create_model()` and then only use
the methods below, so a plugin that implements them is a drop-in (reference train_larva.py:69-74,
validate.py:69-72, get_sr.py:48-51, runtime.py:43-46).
"""


def create_model():
    return BaseModel()


class BaseModel:
    def __init__(self):
        self.global_step = 0
        self.loss_dict = {}

    def parse_args(self, args):
        """Consume this model's flags from `args` (a list of strings); return (namespace, leftover_args)."""
        raise NotImplementedError

    def prepare(self, is_training, scales, global_step=0):
        """Build the network (and, when training, loss/optimizer/scheduler).  Must precede every other call."""
        raise NotImplementedError

    def save(self, base_path):
        """Write a checkpoint of the current weights into directory `base_path`."""
        raise NotImplementedError

    def restore(self, ckpt_path, target=None):
        """Load weights from the checkpoint file `ckpt_path`."""
        raise NotImplementedError

    def get_model(self):
        """Return the underlying torch.nn.Module (may be None)."""
        raise NotImplementedError

    def get_next_train_scale(self):
        """Pick the scale for the next training step."""
        raise NotImplementedError

    def train_step(self, input_list, scale, truth_list, summary=None):
        """Run one optimisation step on a batch; returns a representative loss value."""
        raise NotImplementedError

    def upscale(self, input_list, scale):
        """Super-resolve a list of CHW images without training; returns an NCHW float32 numpy array."""
        raise NotImplementedError
