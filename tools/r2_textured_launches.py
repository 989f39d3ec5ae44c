"""ncu target: one full-size C3 render at 16 spp with a checkerboard Kd on the ground quad and the plastic sphere (the KM_TEX
instantiations of the matte / plastic shade kernels), preceded by one render of the plain scene for comparison.
Usage (GPU box): ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_shade --csv --log-file out.csv python tools/r2_textured_launches.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package()
pkg.init(0)
from pbrt_v3_rs_b200 import workloads as wl

sd = wl.scene_c3(spp=16)
pkg.PathIntegrator(sd).render()
sd = wl.scene_c3(spp=16)
sd.materials[4]["Kd"] = ("texture", sd.add_spectrum_texture("checkerboard", uscale=48.0, vscale=48.0, tex1=(0.1, 0.1, 0.1), tex2=(0.8, 0.8, 0.8)))
sd.materials[1]["Kd"] = ("texture", sd.add_spectrum_texture("checkerboard", uscale=16.0, vscale=16.0, tex1=(0.9, 0.2, 0.1), tex2=(0.1, 0.3, 0.9)))
pkg.PathIntegrator(sd).render()
