"""How wide a shade pass is before and after the block-wide sort of csrc/device/sorted.inl, from the CPU simulation of
the kernel (trace_on_host.render_sorted): live lanes per (warp, hit class) pair.  Test-tier tooling, no GPU.

    python tests/native/sorted_purity.py
"""
import sys
import os
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(os.path.dirname(HERE)), os.path.dirname(HERE), HERE]
import numpy as np
import raytracinginrust_b200 as rt, trace_on_host as toh
from util import host_scene
for name,(W,H,spp) in {"cornell":(64,64,16),"cornell_smoke":(48,48,12),"random":(64,64,8),"final":(48,48,8),"mesh":(64,36,8)}.items():
    hs=host_scene(rt,name); comp=toh.CompiledOnHost(hs.scene_desc)
    opts=rt.render_opts(seed=8, integrator=hs.integrator)
    for block in (128,256):
        img,st=comp.render_sorted(hs.camera,W,H,spp,100,opts,n_chunks=2,n_blocks=4,block=block,purity=True)
        print("%-14s block %3d: live lanes per warp-segment %.1f; width of a shade pass: unsorted %.1f lanes -> sorted %.1f lanes (passes per warp %.2f -> %.2f)" % (
            name, block, st["live_lanes"]/st["warp_segments"], st["live_lanes"]/st["passes_unsorted"], st["live_lanes"]/st["passes_sorted"],
            st["passes_unsorted"]/st["warp_segments"], st["passes_sorted"]/st["warp_segments"]))
