"""GPU SAH builder (csrc/bvh_build.cu) vs the host builder and the oracle's restatement of BVHAccel::new
(accelerators/src/bvh/mod.rs:43-153, sah.rs:26-367): LinearBVHNode bytes and ordered_prims must be identical."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _meshes(wl):
    return {
        "sphere": lambda: wl.displaced_sphere(60, 30),
        "sphere_100k": lambda: wl.displaced_sphere(250, 200),
        "soup": lambda: wl.triangle_soup(5000),
        "soup_200k": lambda: wl.triangle_soup(200000),
        "tiny": lambda: wl.ground_quad(),
        "one": lambda: wl.ground_quad()[:1],
        "n33": lambda: wl.triangle_soup(33),
        "n32": lambda: wl.triangle_soup(32),
        "coincident": lambda: np.tile(wl.ground_quad()[:1], (37, 1)),
        "coincident_big": lambda: np.concatenate([np.tile(wl.ground_quad()[:1], (300, 1)), wl.triangle_soup(700)]),
    }


@pytest.mark.parametrize("max_prims", [1, 4, 255])
@pytest.mark.parametrize("mesh", ["sphere", "sphere_100k", "soup", "soup_200k", "tiny", "one", "n33", "n32", "coincident", "coincident_big"])
def test_gpu_builder_matches_host_and_oracle(pkg, gpu, oracle, mesh, max_prims):
    from pbrt_v3_rs_b200 import workloads as wl
    tv = _meshes(wl)[mesh]()
    pb = pkg.triangle_bounds(tv)
    n0, o0 = pkg.build_bvh_sah(pb, max_prims, where="gpu")
    n1, o1 = pkg.build_bvh_sah(pb, max_prims, where="host")
    assert len(n0) == len(n1)
    assert np.array_equal(o0, o1), "ordered_prims differ"
    assert n0.tobytes() == n1.tobytes(), "LinearBVHNode arrays differ"
    if len(tv) <= 20000:
        n2, o2 = oracle.build_bvh_sah(pb, max_prims)
        assert n0.tobytes() == n2.tobytes() and np.array_equal(o0, o2)


def test_gpu_builder_sign_of_zero(pkg, gpu, oracle):
    """Boxes with +0 / -0 coordinates: the builder reproduces the reference's order-dependent zero signs."""
    rng = np.random.default_rng(11)
    n = 4000
    lo = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    hi = lo + rng.uniform(0.01, 0.2, (n, 3)).astype(np.float32)
    zero = rng.integers(0, 2, (n, 3)).astype(bool)
    lo[zero] = np.where(rng.integers(0, 2, zero.sum()) == 1, np.float32(0.0), np.float32(-0.0))
    hi = np.maximum(hi, lo + np.float32(0.01))
    pb = np.concatenate([lo, hi], axis=1).astype(np.float32)
    for max_prims in (1, 4):
        n0, o0 = pkg.build_bvh_sah(pb, max_prims, where="gpu")
        n2, o2 = oracle.build_bvh_sah(pb, max_prims)
        assert np.array_equal(o0, o2)
        assert n0.tobytes() == n2.tobytes()


def test_gpu_builder_empty(pkg, gpu):
    n, o = pkg.build_bvh_sah(np.zeros((0, 6), np.float32), where="gpu")
    assert len(n) == 0


def test_gpu_triangle_bounds_and_device_build(pkg, gpu):
    """Device-pointer entry points: triangle bounds + build without leaving the GPU."""
    import ctypes as C
    import torch
    from pbrt_v3_rs_b200 import workloads as wl
    tv = wl.displaced_sphere(120, 80)
    n = tv.shape[0]
    d_tv = torch.from_numpy(tv).cuda()
    d_pb = torch.empty((n, 6), dtype=torch.float32, device="cuda")
    L = pkg.lib()
    st = torch.cuda.current_stream().cuda_stream
    assert L.b200pt_triangle_bounds_device(d_tv.data_ptr(), n, d_pb.data_ptr(), st) == 0
    torch.cuda.synchronize()
    assert np.array_equal(d_pb.cpu().numpy(), pkg.triangle_bounds(tv))
    d_nodes = torch.empty((2 * n, 32), dtype=torch.uint8, device="cuda")
    d_ord = torch.empty(n, dtype=torch.int32, device="cuda")
    nn = C.c_int64(0)
    assert L.b200pt_bvh_build_sah_device(d_pb.data_ptr(), n, 4, d_nodes.data_ptr(), C.byref(nn), d_ord.data_ptr(), st) == 0
    n1, o1 = pkg.build_bvh_sah(pkg.triangle_bounds(tv), 4)
    assert nn.value == len(n1)
    assert d_nodes[:nn.value].cpu().numpy().tobytes() == n1.tobytes()
    assert np.array_equal(d_ord.cpu().numpy().view(np.uint32), o1)


def test_full_size_c2_build_is_identical_and_traces(pkg, gpu):
    """1 M triangles (BASELINE config C2): same bytes as the host builder."""
    import time
    from pbrt_v3_rs_b200 import workloads as wl
    tv = wl.c2_mesh(wl.C2_FULL)
    pb = pkg.triangle_bounds(tv)
    pkg.build_bvh_sah(pb[:1000], 4, where="gpu")  # warm-up (context, allocator)
    t0 = time.perf_counter()
    n0, o0 = pkg.build_bvh_sah(pb, 4, where="gpu")
    t1 = time.perf_counter()
    n1, o1 = pkg.build_bvh_sah(pb, 4, where="host")
    t2 = time.perf_counter()
    print("build 1M tris: gpu %.1f ms (host buffers in/out), host %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
    assert np.array_equal(o0, o1)
    assert n0.tobytes() == n1.tobytes()


@pytest.mark.parametrize("max_prims", [1, 4, 255])
@pytest.mark.parametrize("mesh", ["sphere", "sphere_100k", "soup", "soup_200k", "tiny", "one", "n33", "coincident_big"])
def test_gpu_hlbvh_builder_matches_host(pkg, gpu, oracle, mesh, max_prims):
    """SplitMethod::HLBVH on the GPU (Morton codes, stable 5 x 6-bit radix sort, per-treelet LBVH emission) returns the
    host builder's bytes, which the CPU suite pins to the oracle's restatement of hlbvh.rs."""
    import ctypes as C
    import torch
    from pbrt_v3_rs_b200 import workloads as wl
    tv = _meshes(wl)[mesh]()
    pb = pkg.triangle_bounds(tv)
    n = pb.shape[0]
    d_pb = torch.from_numpy(pb).cuda()
    d_nodes = torch.zeros((2 * n, 32), dtype=torch.uint8, device="cuda")
    d_ord = torch.zeros(n, dtype=torch.int32, device="cuda")
    nn = C.c_int64(0)
    rc = pkg.lib().b200pt_bvh_build_hlbvh_device(d_pb.data_ptr(), n, max_prims, d_nodes.data_ptr(), C.byref(nn), d_ord.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0, pkg.lib().b200pt_last_error()
    n1, o1 = pkg.build_bvh_hlbvh(pb, max_prims)
    assert nn.value == len(n1)
    assert np.array_equal(d_ord.cpu().numpy().view(np.uint32), o1), "ordered_prims differ"
    assert d_nodes[:nn.value].cpu().numpy().tobytes() == n1.tobytes(), "LinearBVHNode arrays differ"
    if n <= 20000:
        n2, o2 = oracle.build_bvh_hlbvh(pb, max_prims)
        assert n1.tobytes() == n2.tobytes() and np.array_equal(o1, o2)


@pytest.mark.parametrize("mesh,max_prims", [("sphere_100k", 4), ("soup", 1), ("tiny", 4), ("one", 4), ("coincident", 4)])
def test_accelerator_created_on_the_device_is_the_same_accelerator(pkg, gpu, oracle, mesh, max_prims):
    """b200pt_accel_create_device: triangles in HBM -> bounds -> SAH build -> traversal records without a host round trip.
    Same LinearBVHNode bytes and ordered_prims as the host builder, and bit-identical hits through the default kernels."""
    import torch
    from pbrt_v3_rs_b200 import workloads as wl
    tv = _meshes(wl)[mesh]()
    d_tv = torch.from_numpy(np.ascontiguousarray(tv, dtype=np.float32)).cuda()
    acc = pkg.BVHAccel.from_device_triangles(d_tv.data_ptr(), tv.shape[0], max_prims, stream=torch.cuda.current_stream().cuda_stream, download=True)
    n1, o1 = pkg.build_bvh_sah(pkg.triangle_bounds(tv), max_prims, where="host")
    assert acc.nodes.tobytes() == n1.tobytes() and np.array_equal(acc.ordered_prims, o1)
    assert np.array_equal(acc.world_bound(), n1[0]["bounds"])
    ref = pkg.BVHAccel(tv, n1, o1)
    rays = wl.primary_rays(128, 64, eye=(0.0, 0.3, -3.5)) if mesh not in ("tiny", "one", "coincident") else wl.primary_rays(64, 32, eye=(0.0, 4.0, -0.5))
    if mesh in ("tiny", "one", "coincident"):
        rays["d"] = np.array([0.02, -1.0, 0.03], np.float32) / np.float32(np.linalg.norm([0.02, -1.0, 0.03]))
        rays["o"] = np.stack([np.linspace(-7, 7, rays.shape[0]), np.full(rays.shape[0], 4.0), np.linspace(-7, 7, rays.shape[0])[::-1]], axis=1).astype(np.float32)
    h0, h1 = acc.intersect_batch(rays), ref.intersect_batch(rays)
    assert h0.tobytes() == h1.tobytes()
    assert (h0["prim"] != pkg.MISS).any()
    sh = wl.shadow_rays(rays)
    assert np.array_equal(acc.occluded_batch(sh), ref.occluded_batch(sh))


def test_hlbvh_host_buffer_entry_point(pkg, gpu):
    from pbrt_v3_rs_b200 import workloads as wl
    pb = pkg.triangle_bounds(wl.displaced_sphere(100, 60))
    n0, o0 = pkg.build_bvh_hlbvh(pb, 4, where="gpu")
    n1, o1 = pkg.build_bvh_hlbvh(pb, 4, where="host")
    assert n0.tobytes() == n1.tobytes() and np.array_equal(o0, o1)
