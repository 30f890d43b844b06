"""Data-loader plugin contract (mirrors reference dataloaders/base.py:9-110).

Entry scripts obtain a loader with `importlib.import_module('dataloaders.' + name).create_loader()`; images are CHW
numpy arrays, RGB, value range 0..255 (reference dataloaders/div2k_val_loader.py:133-137).
"""


def create_loader():
    return BaseLoader()


class BaseLoader:
    def __init__(self):
        self.is_threaded = False

    def parse_args(self, args):
        """Consume this loader's flags; return (namespace, leftover_args)."""
        raise NotImplementedError

    def prepare(self, scales):
        """Get ready to serve data for the given list of scales."""
        raise NotImplementedError

    def get_num_images(self):
        raise NotImplementedError

    def get_patch_batch(self, batch_size, scale, input_patch_size):
        """Return (input_list, truth_list): `batch_size` LR patches of `input_patch_size` and their HR truths."""
        raise NotImplementedError

    def get_random_image_patch_pair(self, scale, input_patch_size):
        raise NotImplementedError

    def get_image_patch_pair(self, image_index, scale, input_patch_size):
        raise NotImplementedError

    def get_image_pair(self, image_index, scale):
        """Return (input_image, truth_image, image_name) for a whole validation image."""
        raise NotImplementedError

    def start_training_queue_runner(self, batch_size, input_patch_size):
        raise NotImplementedError

    def stop_queue_runners(self):
        raise NotImplementedError

    def get_queue_data(self, scale):
        raise NotImplementedError
