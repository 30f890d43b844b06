"""Training entry point for LarvaNetV2 (drop-in for reference train_larvaV2.py: train_larva.py plus
`--steps_per_epoch` / `model.steps_per_epoch` epoch bookkeeping)."""
from train_larva import main as _main


def main(argv=None):
    return _main(argv, epoch_bookkeeping=True)


if __name__ == '__main__':
    main()
