"""Inference entry point: directory of PNGs -> x4 super-resolved PNGs (same command line as the reference's get_sr.py).
Implementation: larvanet_b200/entrypoints.py.

    python get_sr.py --model=LarvaNet --num_modules=4 --num_blocks=4,4,4,4 --restore_path=ckpt.pth \
        --input_path=lr/ --output_path=sr/
"""
from larvanet_b200.entrypoints import get_sr_main as main

if __name__ == '__main__':
    main()
