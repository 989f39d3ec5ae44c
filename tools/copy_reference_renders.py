"""Copies the reference's own committed renders of the ten shipped scenes that lie entirely on this path (triangle
meshes, matte incl. the checkerboard Kd texture, point / spot / distant / infinite light, Whitted, Halton, box filter) into
tests/golden/ref_renders/:

  renders/lights/point.png                 <- scenes/lights/point.pbrt
  renders/lights/infinite-no-map.png       <- scenes/lights/infinite-no-map.pbrt
  renders/shapes/triangles-alpha-mask.png  <- scenes/shapes/triangles-alpha-mask.pbrt
  renders/lights/distant.png               <- scenes/lights/distant.pbrt        (distant light, "blackbody L")
  renders/objects/instances.png            <- scenes/objects/instances.pbrt     (ten ObjectInstances of a cube: the two-level BVH)
  renders/cameras/perspective.png          <- scenes/cameras/perspective.pbrt   (infinite + distant light)
  renders/lights/spot.png                  <- scenes/lights/spot.pbrt           (spot light; its "conedelta" is not a parameter the reference reads)
  renders/cameras/orthographic.png         <- scenes/cameras/orthographic.pbrt  (orthographic camera, the world under a Scale)
  renders/cameras/environment.png          <- scenes/cameras/environment.pbrt   (environment camera, 800 x 400, ten cubes around it)
  renders/lights/goniometric.png           <- scenes/lights/goniometric.pbrt    (goniometric light: a 1572 x 790 image through MIPMap::new's resampling)

They are OUTPUTS of the reference (8-bit sRGB PNGs written by core/src/image_io.rs), i.e. golden vectors: the only
artefacts in the reference tree that were produced by executing it.  tests/test_reference_renders.py renders the same
scenes with the oracle (CPU) and with the CUDA path (GPU) and compares after the reference's gamma / 8-bit encode."""
import os
import shutil

SRC = "/root/reference/renders"
DST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ref_renders")

if __name__ == "__main__":
    os.makedirs(DST, exist_ok=True)
    for sub, name in (("lights", "point"), ("lights", "infinite-no-map"), ("shapes", "triangles-alpha-mask"), ("lights", "distant"),
                      ("objects", "instances"), ("cameras", "perspective"), ("lights", "spot"), ("cameras", "orthographic"), ("cameras", "environment"), ("lights", "goniometric")):
        shutil.copyfile(os.path.join(SRC, sub, name + ".png"), os.path.join(DST, name + ".png"))
        print("copied", name)
    # the goniometric light's own input image (scene data, read by scenes/lights/goniometric.pbrt)
    shutil.copyfile("/root/reference/scenes/images/goniometric-upward-downward.png", os.path.join(DST, "goniometric-upward-downward.png"))
