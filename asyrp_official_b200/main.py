"""CLI of the editing / inference path: the flags of the reference's main.py that reach `Asyrp.run_test`
(SURVEY.md Appendix E), same names and defaults (main.py:12-229), same experiment-directory naming (:235).

    python -m asyrp_official_b200.main --run_test --config celeba.yml --exp ./runs/smiling --edit_attr smiling \
        --train_delta_block --get_h_num 1 --load_random_noise --user_defined_t_edit 500 --user_defined_t_addnoise 200 \
        --manual_checkpoint_name smiling_LC_CelebA_HQ_t999_ninv40_ngen40_0.pth --model_path celeba_hq.ckpt \
        --n_test_img 32 --bs_train 16 --n_test_step 40 --n_train_step 40

Multi-GPU: launch with torchrun (one process per GPU); image batches are sharded across ranks.
"""
import argparse
import logging
import os
import sys
import traceback

import numpy as np
import torch

from .configs import load_config


def build_parser():
    p = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    sa = dict(action='store_true')
    p.add_argument('--sh_file_name', type=str, default='script.sh')
    p.add_argument('--user_defined_t_edit', type=int)
    p.add_argument('--user_defined_t_addnoise', type=int)
    p.add_argument('--lpips_edit_th', type=float, default=0.33)
    p.add_argument('--lpips_addnoise_th', type=float, default=0.1)
    p.add_argument('--add_noise_from_xt', **sa)
    p.add_argument('--origin_process_addnoise', **sa)
    p.add_argument('--run_test', **sa)
    p.add_argument('--train_delta_block', **sa)
    p.add_argument('--train_delta_h', **sa)
    p.add_argument('--ignore_timesteps', **sa)
    p.add_argument('--use_x0_tensor', **sa)
    p.add_argument('--save_x0', **sa)
    p.add_argument('--save_x_origin', **sa)
    p.add_argument('--load_random_noise', **sa)
    p.add_argument('--saved_random_noise', **sa)
    p.add_argument('--delta_interpolation', **sa)
    p.add_argument('--max_delta', type=float, default=1.0)
    p.add_argument('--min_delta', type=float, default=0.0)
    p.add_argument('--num_delta', type=int, default=5)
    p.add_argument('--hs_coeff_delta_h', type=float, default=1.0)
    p.add_argument('--hs_coeff_origin_h', type=float, default=1.0)
    p.add_argument('--target_image_id', type=str)
    p.add_argument('--start_image_id', type=int, default=0)
    p.add_argument('--save_process_origin', **sa)
    p.add_argument('--save_process_delta_h', **sa)
    p.add_argument('--num_mean_of_delta_hs', type=int, default=0)
    p.add_argument('--multiple_attr', type=str, default='')
    p.add_argument('--multiple_hs_coeff', type=str, default='')
    p.add_argument('--manual_checkpoint_name', type=str, default="")
    p.add_argument('--choose_checkpoint_num', type=str, default='')
    p.add_argument('--load_from_checkpoint', type=str)
    p.add_argument('--pass_editing', **sa)
    p.add_argument('--warigari', type=float, default=0.0)
    p.add_argument('--config', type=str, required=True)
    p.add_argument('--seed', type=int, default=1234)
    p.add_argument('--exp', type=str, default='./runs/')
    p.add_argument('--verbose', type=str, default='info')
    p.add_argument('--ni', type=int, default=1)
    p.add_argument('--edit_attr', type=str, default=None)
    p.add_argument('--t_0', type=int, default=999)
    p.add_argument('--n_inv_step', type=int, default=40)
    p.add_argument('--n_train_step', type=int, default=6)
    p.add_argument('--n_test_step', type=int, default=40)
    p.add_argument('--sample_type', type=str, default='ddim')
    p.add_argument('--do_train', type=int, default=1)
    p.add_argument('--do_test', type=int, default=1)
    p.add_argument('--bs_train', type=int, default=1)
    p.add_argument('--n_train_img', type=int, default=50)
    p.add_argument('--n_test_img', type=int, default=10)
    p.add_argument('--model_path', type=str, default=None)
    p.add_argument('--get_h_num', type=int, default=0)
    p.add_argument('--re_precompute', **sa)
    p.add_argument('--custom_train_dataset_dir', type=str, default="./custom/train")
    p.add_argument('--custom_test_dataset_dir', type=str, default="./custom/test")
    p.add_argument('--dt_lambda', type=float, default=1.0)
    p.add_argument('--dt_end', type=int, default=950)
    p.add_argument('--n_iter', type=int, default=1)
    # additions of this build (documented in INTEGRATION.md)
    p.add_argument('--synthetic_weights', **sa, help='seeded random UNet weights instead of --model_path')
    p.add_argument('--clip_cosine', type=float, help='text-direction cosine CLIP would give (no CLIP offline)')
    p.add_argument('--lpips_table_dir', type=str, help='directory with <category>_LPIPS_distance_{x,x0_t}.tsv')
    p.add_argument('--checkpoint_dir', type=str, default='checkpoint')
    return p


def parse_args_and_config(argv=None):
    args = build_parser().parse_args(argv)
    config = load_config(args.config)
    args.exp = args.exp + f'_LC_{config.data.category}_t{args.t_0}_ninv{args.n_inv_step}_ngen{args.n_train_step}'
    level = getattr(logging, args.verbose.upper(), None)
    if not isinstance(level, int):
        raise ValueError('level {} not supported'.format(args.verbose))
    logging.basicConfig(level=level, format='%(levelname)s - %(filename)s - %(asctime)s - %(message)s')
    for d in ('checkpoint', 'checkpoint_latent', 'precomputed', 'runs', args.exp):
        os.makedirs(d, exist_ok=True)
    args.test_image_folder = os.path.join(args.exp, 'test_images', str(args.n_test_step))
    args.image_folder = os.path.join(args.exp, 'image_samples')
    os.makedirs(args.test_image_folder, exist_ok=True)
    os.makedirs(args.image_folder, exist_ok=True)
    torch.manual_seed(args.seed)
    np.random.seed(args.seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(args.seed)
    return args, config


def main(argv=None):
    args, config = parse_args_and_config(argv)
    from .diffusion_latent import Asyrp
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    runner = Asyrp(args, config)
    try:
        if args.run_test:
            runner.run_test()
        else:
            print('Choose one mode! (this build implements --run_test)')
            raise ValueError
    except Exception:
        logging.error(traceback.format_exc())
        return 1
    return 0


if __name__ == '__main__':
    sys.exit(main())
