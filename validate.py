"""Validation entry point (same command line as the reference's validate.py) and the PSNR helpers the model plugins
import from here.  Implementation: larvanet_b200/entrypoints.py.

    python validate.py --model=LarvaNet --num_modules=4 --num_blocks=4,4,4,4 --restore_path=ckpt.pth \
        [--dataloader=synthetic_val_loader] [--save_path=out/] [--chop_forward]
"""
from larvanet_b200.entrypoints import (_fit_truth_image_size, _image_psnr, _image_to_uint8, _save_image,  # noqa: F401
                                       validate_main as main)

if __name__ == '__main__':
    main()
