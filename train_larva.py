"""Training entry point for LarvaNet (same command line as the reference's train_larva.py; `train_larvaV2.py` adds the
epoch bookkeeping).  Implementation: larvanet_b200/entrypoints.py.

    python train_larva.py --model=LarvaNet --num_modules=4 --num_blocks=4,4,4,4 --batch_size=16 --input_patch_size=48 \
        --train_path=/tmp/larva --max_steps=100 --sleep_ratio=0 [--dataloader=synthetic_loader]
"""
from larvanet_b200.entrypoints import train_main as main

if __name__ == '__main__':
    main()
