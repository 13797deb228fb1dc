"""-m gpu: ITMMeshingEngine::MeshScene + ITMMesh::WriteSTL/WriteOBJ against the reference CPU meshing
engine (ITMMeshingEngine_CPU.cpp:19-58).  The triangle array must be BIT EXACT and in the same order (the CUDA path
reproduces the serial loop's append order with an ordered scan), hence the files are byte identical."""
import ctypes as C
import filecmp

import numpy as np
import pytest

import parity
from infinitam_b200 import capi, synth
from oracle import adapter, ref

pytestmark = pytest.mark.gpu

needs_ref = pytest.mark.skipif(not ref.available("parity"), reason="oracle/_ref not built (needs /root/reference at build time)")


def _fused_pair(w, h, n, flavour="parity"):
    seq = synth.sequence(n, w, h)
    o = ref.RefEngine(w, h, flavour=flavour)
    eng = parity.make_cuda_engine(o)
    for k in range(n):
        o.process_frame(seq[k])
    eng.UploadDepth(seq[-1])
    parity.push_scene(o, eng)   # identical scenes on both sides
    return o, eng


@needs_ref
@pytest.mark.parametrize("flavour", ["parity", "rgb"])
def test_mesh_scene_bit_exact(tmp_path, flavour):
    if not ref.available(flavour):
        pytest.skip("oracle/_ref flavour %s not built" % flavour)
    o, eng = _fused_pair(320, 240, 4, flavour)
    tri_ref = o.mesh_scene()
    tri_gpu = eng.UpdateMesh()
    assert len(tri_ref) > 10000, "the scene should produce a real mesh"
    assert tri_gpu.shape == tri_ref.shape, "noTotalTriangles: gpu %d ref %d" % (len(tri_gpu), len(tri_ref))
    assert np.array_equal(tri_gpu.view(np.uint32), tri_ref.view(np.uint32)), "triangles differ"
    # ITMMesh::WriteSTL / SaveSceneToMesh and WriteOBJ: byte-identical files
    o.write_stl(tmp_path / "ref.stl")
    eng.SaveSceneToMesh(tmp_path / "gpu.stl")
    assert filecmp.cmp(tmp_path / "ref.stl", tmp_path / "gpu.stl", shallow=False)
    k = 5000
    o.no_total = None
    lib = capi.load()
    capi.check(lib.itm_b200_write_obj(str(tmp_path / "gpu.obj").encode(), tri_gpu.ctypes.data, len(tri_gpu)))
    o.write_obj(tmp_path / "ref.obj")
    assert filecmp.cmp(tmp_path / "ref.obj", tmp_path / "gpu.obj", shallow=False)
    eng.close(); o.close()


@needs_ref
def test_mesh_scene_capacity_rule():
    """a mesh smaller than the scene's triangle count: the write index stops at noMaxTriangles - 1, so that slot holds the
    last triangle emitted (ITMMeshingEngine_CPU.cpp:51) - Layer A with a caller-owned, deliberately small mesh buffer"""
    import torch
    o, eng = _fused_pair(320, 240, 2)
    tri_ref = np.array(o.mesh_scene(), copy=True)
    n_max = len(tri_ref) // 3
    lib = capi.load()
    ctx = C.c_void_p()
    capi.check(lib.itm_b200_ctx_create(C.byref(eng.params), None, C.byref(ctx)))
    scene = capi.Scene()
    scene.voxel_blocks_dev, _ = eng.buffer_info(capi.BUF_VOXELS)
    scene.hash_entries_dev, _ = eng.buffer_info(capi.BUF_HASH)
    buf = torch.full((n_max + 8, 9), 7.0, dtype=torch.float32, device="cuda")   # 8 guard rows behind the mesh
    n = C.c_uint()
    capi.check(lib.itm_b200_mesh_scene(ctx, C.byref(scene), C.c_void_p(buf.data_ptr()), n_max, C.byref(n)))
    out = buf.cpu().numpy()
    assert n.value == n_max - 1
    assert np.array_equal(out[:n_max - 1], tri_ref[:n_max - 1])
    assert np.array_equal(out[n_max - 1], tri_ref[-1])
    assert (out[n_max:] == 7.0).all(), "wrote past the mesh"
    lib.itm_b200_ctx_destroy(ctx)
    eng.close(); o.close()


@pytest.mark.skipif(not (adapter.available() and ref.available("parity")), reason="oracle/_ref not built")
def test_adapter_save_scene_to_mesh(tmp_path):
    """ITMMeshingEngine_B200 behind the reference's own ITMMesh (CUDA memory) and its own WriteSTL"""
    w, h = 320, 240
    seq = synth.sequence(1, w, h)
    o = ref.RefEngine(w, h)
    a = adapter.AdapterEngine(w, h, intr=o.intr)
    o.process_frame(seq[0])   # frame 0 runs at the identity pose on both sides: identical scenes
    a.process_frame(seq[0])
    o.mesh_scene()
    o.write_stl(tmp_path / "ref.stl")
    n = a.save_scene_to_mesh(tmp_path / "adp.stl")
    assert n == o.no_total_triangles and n > 1000
    assert filecmp.cmp(tmp_path / "ref.stl", tmp_path / "adp.stl", shallow=False)
    a.close(); o.close()


def test_empty_scene_rows_8f(tmp_path):
    """edge cases of the 8f rows on a scene that holds nothing: MeshScene yields no triangle (an 84-byte STL: header +
    count), the free-view rendering is black with an empty visible list, and an all-invalid depth frame leaves it so"""
    w, h = 160, 120
    p = capi.default_params(w, h)
    p.use_approximate_raycast = 1
    from infinitam_b200.engines import ITMMainEngine
    eng = ITMMainEngine(p)
    assert len(eng.UpdateMesh()) == 0
    eng.ProcessFrame(None, np.zeros((h, w), np.int16))   # no valid pixel
    eng.ProcessFrame(None, np.full((h, w), -5, np.int16))
    _, _, st = eng.get_state()
    assert int(st[0]) == 0 and int(st[1]) == p.sdf_local_block_num - 1, "an empty frame must not allocate"
    assert len(eng.UpdateMesh()) == 0
    eng.SaveSceneToMesh(tmp_path / "empty.stl")
    assert (tmp_path / "empty.stl").stat().st_size == 84
    K = synth.intrinsics_for(w, h)
    M = np.eye(4, dtype=np.float32).reshape(16)
    for t in (capi.IMAGE_FREECAMERA_SHADED, capi.IMAGE_FREECAMERA_COLOUR_FROM_NORMAL, capi.IMAGE_FREECAMERA_COLOUR_FROM_VOLUME, capi.IMAGE_SCENERAYCAST):
        assert not eng.GetImage(t, M, K).any()
    eng.close()
