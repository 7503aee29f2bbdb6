"""Run under torchrun (one rank per GPU): the imageset-sharded multi-GPU LM loop must follow the
single-GPU loop (same accept sequence, costs to 1e-9 relative, same final state)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from camera_calibration_b200 import api, cabi, distributed, synthetic
    rank, world, local = distributed.env_rank_world()
    dist = distributed.init_process_group("nccl")
    torch.cuda.set_device(local)
    worst = 0.0
    # second sweep with the structured (grouped) Schur contraction forced on
    if os.environ.get("MGPU_FORCE_GROUPED"):
        os.environ["B200BA_GROUPED"] = "1"
        os.environ["B200BA_GROUP_BLOCKS"] = "7"
    # third sweep: narrow column blocks, so that the block-cyclic Cholesky (panel broadcasts, look-ahead)
    # spans many panels on the small problems
    if os.environ.get("MGPU_SMALL_PANELS"):
        os.environ["B200BA_DENSE_NB"] = "128"
    for cfg, kw in ((2, dict(n_imagesets=12, lattice=(12, 10), image_size=(410, 290))),
                    (4, dict(n_imagesets=10, lattice=(10, 8), image_size=(410, 290))),
                    (1, dict(n_imagesets=8, lattice=(10, 10)))):
        sp = synthetic.make_problem(cfg, **kw)
        opt = cabi.default_options(max_iteration_count=6)
        adj, idx = distributed.make_sharded_adjuster(sp.problem)
        st = distributed.shard_state(sp.init_state, idx)
        rep = adj.optimize_host(st, opt)
        # single-GPU reference run of the same problem on this rank's device
        with api.BundleAdjuster(sp.problem, local) as ref:
            rst = sp.init_state.copy()
            rrep = ref.optimize_host(rst, opt)
        gc, gl, ga = rep.trace()
        rc, rl, ra = rrep.trace()
        assert ga == ra, (cfg, ga, ra)
        assert np.allclose(gc, rc, rtol=1e-9), (cfg, gc, rc)
        assert abs(rep.rmse - rrep.rmse) < 1e-9
        assert rep.n_valid == rrep.n_valid
        d = max(np.abs(st.points - rst.points).max(), np.abs(st.rig_tr_global - rst.rig_tr_global).max(),
                max(np.abs(a - b).max() for a, b in zip(st.intrinsics, rst.intrinsics)))
        assert d < 1e-9, (cfg, d)
        assert np.abs(st.last_projection - rst.last_projection[idx]).max() < 1e-9
        worst = max(worst, d)
        adj.close()
        dist.barrier()
    if rank == 0:
        print(f"MGPU_OK world={world} max_state_diff={worst:.3e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
