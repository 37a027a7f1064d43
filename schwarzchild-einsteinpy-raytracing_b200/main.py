"""Render entry point — same flow as the reference's main.py:22-157: parse flags, build the scene,
flat comparison render, curved render, save images, diagnostic plots (when matplotlib is available),
photon summary from photon_data.csv.  Geometrized units, G = c = 1.
"""
import logging
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from config import parse_args  # noqa: E402
from simulation.background import save_no_gravity_image_with_background  # noqa: E402
from simulation.blackhole import BlackHole, Observer  # noqa: E402
from simulation.raytracing import run_manual_simulation, _save_png  # noqa: E402
from simulation.utils import _apply_relative_offsets  # noqa: E402

logging.basicConfig(level=logging.INFO, format='%(asctime)s %(levelname)s: %(message)s')


def _init_distributed():
    """Under torchrun (WORLD_SIZE > 1): one process per GPU; rows are sharded by simulation.raytracing / schw_b200.multigpu."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return dist.get_rank()


def main(argv=None):
    args = parse_args(argv)
    rank = _init_distributed()
    image_size = (args.height or args.size, args.width or args.size)
    bh = BlackHole(mass=args.bh_mass)
    observer = Observer(position=np.array([args.observer_distance, 0, 0]), fov=np.radians(args.fov), image_size=image_size)
    pct, pcp = _apply_relative_offsets(args.bg_patch_center_theta, args.bg_patch_center_phi,
                                       args.bg_patch_center_theta_relobs, args.bg_patch_center_phi_relobs)
    patch = dict(patch_center_theta=pct, patch_center_phi=pcp,
                 patch_size_theta=np.deg2rad(args.bg_patch_size_theta), patch_size_phi=np.deg2rad(args.bg_patch_size_phi),
                 flip_theta=args.bg_flip_theta, flip_phi=args.bg_flip_phi)

    logging.info("Saving no-gravity image using background...")
    flat_trajectories = None
    if not args.no_flat_trajectories and rank == 0:
        flat_trajectories = save_no_gravity_image_with_background(
            observer, args.background, 'images/no_gravity.png', boundary_radius=args.boundary_radius,
            return_sampled_trajectories=True, n_sampled=10, override_patch_center=False, **patch)

    result = run_manual_simulation(
        bh, observer, real_time=False, update_every=max(1, (image_size[0] * image_size[1]) // 32),
        steps=args.steps, delta=args.delta, omega=args.omega, rtol=args.rtol, atol=args.atol, order=args.order,
        suppress_warnings=args.suppress_warnings, background_path=args.background, use_cuda=args.cuda,
        boundary_radius=args.boundary_radius, n_samples=20, dedup=args.dedup, **patch)
    img, photon_trajectories = result if isinstance(result, tuple) else (result, None)
    if rank != 0:
        return                                  # rank 0 owns the files and the summary
    _save_png('images/manual_output.png', img)
    logging.info("Saved manual_output.png")

    if photon_trajectories is not None:
        # drop the all-zero rows the decimation of zero-padded trajectories leaves behind (main.py:106-114)
        kept = [t[~np.all(t == 0, axis=1)] for t in photon_trajectories]
        photon_trajectories = [t for t in kept if len(t) > 0]
        print(f"Filtered {len(photon_trajectories)} trajectories")

    if not args.no_plots:
        try:
            from visualization.plot import plot_scene_topdown, plot_scene_embedding_3d, plot_scene_closeup_3d
            plot_scene_topdown(bh, observer, image_size, boundary_radius=args.boundary_radius,
                               out_path='images/scene_topdown.png', fov_deg=args.fov, patch_center_theta=pct,
                               patch_size_theta=patch["patch_size_theta"], patch_size_phi=patch["patch_size_phi"],
                               photon_trajectories=photon_trajectories)
            plot_scene_closeup_3d(bh, observer, image_size, out_path='images/scene_closeup_3d.png', fov_deg=args.fov,
                                  photon_trajectories=photon_trajectories)
            plot_scene_embedding_3d(bh, observer, image_size, boundary_radius=args.boundary_radius,
                                    out_path='images/scene_topdown_3d.png', fov_deg=args.fov,
                                    photon_trajectories=photon_trajectories, flat_trajectories=flat_trajectories,
                                    patch_center_theta=pct, patch_center_phi=pcp,
                                    patch_size_theta=patch["patch_size_theta"], patch_size_phi=patch["patch_size_phi"],
                                    override_patch_center=False)
        except ImportError as e:      # the reference's visualization/ package (matplotlib) is out of scope here
            logging.info("Diagnostic plots skipped (%s)", e)

    try:
        import pandas as pd
        df = pd.read_csv('photon_data.csv')
        print("\nPhoton summary:")
        print(f"  Captured by BH: {(df['collision'] == 'bh').sum()}")
        print(f"  Still in domain: {(df['collision'] == 'in_domain').sum()}")
        print(f"  Escaped: {((df['collision'] == 'escape_no_patch') | (df['collision'] == 'escape_bg')).sum()}")
        print(f"  Hit background: {(df['collision'] == 'escape_bg').sum()}")
    except Exception as e:
        print(f"Could not read photon_data.csv for summary: {e}")


if __name__ == "__main__":
    main()
