"""Run the BASELINE.json configurations 2-5 on one GPU (or under torchrun for config 5) and print one JSON line each."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import wasm_pathtracer_b200 as W
from bench import mesh_path

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from wasm_pathtracer_b200.dist import attach
verts = W.parse_obj(open(mesh_path()).read(), True)
which = sys.argv[1:] or ["2", "3", "3m", "4", "5"]
scale = float(os.environ.get("WPT_CFG_SCALE", "1"))   # scale the sample budgets (1 = BASELINE sizes)


def run(name, scene, cam, w, h, bvh, rtype, mode, spp):
    pt = W.PathTracer(w, h, scene, *cam, device=local)
    if scene == W.SCENE_BUNNY:
        pt.store_mesh(1, verts)
    pt.set_config(bvh_kind=bvh, render_type=rtype)
    attach(pt, rank, world)   # band partition + native NCCL plane: accumulator all-gather between adaptive rounds, photon shots split over ranks
    t0 = time.perf_counter()
    if rtype == W.PNEE:
        pt.build_photons()
    pt.synchronize()
    t_ph = time.perf_counter() - t0
    t0 = time.perf_counter()
    if mode == "exact":
        pt.render_exact(spp)
        pt.gather_frame()
    else:
        pt.render_adaptive(int(w * h * spp))
    img = pt.results(0)
    dt = time.perf_counter() - t0
    st = pt.stats()
    tt = torch.tensor([dt, st["rays"] - st["photons_shot"], st["paths"]], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = tt.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX); sm = tt.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dt = float(mx[0]); rays = float(sm[1]); paths = float(sm[2])
    else:
        rays, paths = float(tt[1]), float(tt[2])
    if rank == 0:
        rgb, cnt = pt.accum()
        print(json.dumps({"config": name, "n_gpus": world, "viewport": [w, h], "bvh": bvh, "render_type": rtype, "mode": mode, "spp": spp,
                          "seconds": dt, "photon_warmup_s": t_ph, "Mrays_per_s": rays / dt / 1e6, "Mpaths_per_s": paths / dt / 1e6,
                          "rays": rays, "paths": paths, "spp_min": int(cnt.min()), "spp_max": int(cnt.max()), "mean_rgb": [float(x) for x in (rgb.sum((0, 1)) / max(1, cnt.sum()))],
                          "photons": st["photons_stored"], "photon_shots": st["photons_shot"]}), flush=True)
    pt.detach_nccl()
    pt.close()


for c in which:
    if c == "2": run("2: bunny BVH2 1080p 16spp NormalNEE", W.SCENE_BUNNY, W.CAM_BUNNY, 1920, 1080, 2, W.NORMAL_NEE, "exact", max(1, int(16 * scale)))
    if c == "3": run("3: bunny BVH4 1080p 16spp PNEE (300k photons)", W.SCENE_BUNNY, W.CAM_BUNNY, 1920, 1080, 4, W.PNEE, "exact", max(1, int(16 * scale)))
    if c == "3m": run("3m: museum BVH2 1080p 16spp PNEE (108 lights)", W.SCENE_MUSEUM, W.CAM_MUSEUM, 1920, 1080, 2, W.PNEE, "exact", max(1, int(16 * scale)))
    if c == "4": run("4: museum 4K adaptive 256spp budget NormalNEE", W.SCENE_MUSEUM, W.CAM_MUSEUM, 3840, 2160, 2, W.NORMAL_NEE, "adaptive", max(4, int(256 * scale)))
    if c == "5": run("5: bunny 4K PNEE + adaptive 1024spp budget", W.SCENE_BUNNY, W.CAM_BUNNY, 3840, 2160, 4, W.PNEE, "adaptive", max(4, int(1024 * scale)))
if dist is not None:
    dist.barrier(); dist.destroy_process_group()
