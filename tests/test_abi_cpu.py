"""CPU tests of the drop-in boundary: the C-ABI library loads (no GPU needed) and exports every symbol that
include/claymore_b200.h declares; host-side logic (config, samplers, scene schema)."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "claymore_b200.h")).read()
    return sorted(set(re.findall(r"CB200_API[^;(]*?\b(cb200_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import claymore_b200 as cb
    cb.build_library()
    lib = C.CDLL(cb.lib_path())
    names = _declared_symbols()
    assert len(names) >= 50
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/claymore_b200.h but not exported: {missing}"
    assert lib.cb200_version is not None


def test_binding_signatures_cover_header():
    from claymore_b200 import _capi
    bound = set(_capi._SIGNATURES) | {"cb200_version", "cb200_error_string", "cb200_sim_launch_count", "cb200_default_material"}
    assert set(_declared_symbols()) <= bound, sorted(set(_declared_symbols()) - bound)


def test_struct_layouts_match_header():
    """ctypes mirrors of the by-value structs: sizes as the C compiler lays them out."""
    import subprocess
    import tempfile
    from claymore_b200 import _capi
    prog = '#include <stdio.h>\n#include "claymore_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(cb200_config), sizeof(cb200_particle_buffer), sizeof(cb200_partition), sizeof(cb200_sim_desc), sizeof(cb200_sim_stats));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(prog)
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")])
        sizes = [int(x) for x in subprocess.check_output([os.path.join(d, "t")]).split()]
    assert sizes == [C.sizeof(_capi.Config), C.sizeof(_capi.ParticleBuffer), C.sizeof(_capi.Partition), C.sizeof(_capi.SimDesc), C.sizeof(_capi.SimStats)]


def test_config_derived_quantities():
    import claymore_b200 as cb
    c = cb.Config(domain_bits=9, max_ppc=64)
    assert c.grid_size == 128 and c.ppb == 4096 and abs(c.dx - 1 / 512) < 1e-12


def test_samplers_match_reference_lattice():
    from claymore_b200 import samplers
    dx = 1.0 / 128
    p = samplers.uniform_box(dx, (51, 51, 51), (77, 77, 77))
    assert p.shape == (140608, 3)                      # 26^3 cells x 8 (BASELINE config 1)
    cells = np.round(p / dx).astype(int)
    assert cells.min() == 51 and cells.max() == 76
    frac = p / dx - cells
    assert np.allclose(np.abs(frac), 0.25, atol=1e-4)  # +-0.25 dx around i*dx (GeometrySampler.h:22-29)
    s = samplers.sphere(1.0 / 64, (0.5, 0.5, 0.5), 0.1)
    assert np.all(((s - 0.5) ** 2).sum(1) <= 0.1 ** 2 + 1e-9) and 2000 < len(s) < 12000
    parts = samplers.split_slabs(s, 4)
    assert sum(len(x) for x in parts) == len(s) and max(len(x) for x in parts) - min(len(x) for x in parts) <= 1
    assert all(parts[i][:, 0].max() <= parts[i + 1][:, 0].min() + 1e-9 for i in range(3))


def test_scene_schema_mapping():
    from claymore_b200 import scene
    import claymore_b200 as cb
    assert scene.CONSTITUTIVE == {"jfluid": cb.J_FLUID, "fixed_corotated": cb.FIXED_COROTATED, "sand": cb.SAND, "nacc": cb.NACC}
    cfg = cb.Config(domain_bits=6)
    pos = scene._positions({"file": "box", "offset": [0.25, 0.25, 0.25], "span": [0.125, 0.125, 0.125]}, cfg, ".")
    assert pos.shape == (8 * 8 * 8 * 8, 3)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from claymore_b200 import _capi
    monkeypatch.setattr(_capi, "_lib", None)
    monkeypatch.setattr(_capi, "_LIB", str(tmp_path / "nope.so"))
    with pytest.raises(_capi.CB200Error):
        _capi.lib()


def test_adapter_header_compiles_against_the_reference(tmp_path):
    """include/claymore_b200_adapter.cuh -- the C++ overloads with the reference's kernel argument lists -- must compile against the
    reference's own headers with every functor instantiated for every material (only where /root/reference exists)."""
    import shutil
    import subprocess
    ref = os.environ.get("CLAYMORE_REFERENCE", "/root/reference")
    if not os.path.isdir(ref) or not shutil.which("nvcc") and not os.path.exists("/usr/local/cuda/bin/nvcc"):
        pytest.skip("reference checkout or nvcc not available")
    tu = tmp_path / "adapter_tu.cu"
    tu.write_text(r"""
#include <MnBase/Math/Matrix/Givens.cuh>
namespace mn { namespace math { template<typename T> __host__ __device__ void polar_decomposition(const std::array<T, 4>& a, GivensRotation<T>& r, std::array<T, 4>& s); }}
#include <claymore_b200_adapter.cuh>
using namespace mn;
struct Ctx { cudaStream_t stream_compute() { return nullptr; } };
template<MaterialE M> void every_call(Ctx& cu, ParticleBuffer<M>& pb, ParticleBuffer<M>& next, Partition<1>& part, Partition<1>& prev, GridBuffer& g0, GridBuffer& g1, ParticleArray& pa, int* marks, float* mv) {
    b200::compute_launch(cu, {8, 128}, b200::g2p2g, Duration(1e-4f), Duration(1e-4f), (const ParticleBuffer<M>) pb, next, (const Partition<1>) prev, part, (const GridBuffer) g0, g1);
    b200::compute_launch(cu, {8, 128}, b200::update_buckets, (uint32_t) 8, (const int*) marks, (const ParticleBuffer<M>) pb, next);
    b200::compute_launch(cu, {8, 128}, b200::build_particle_cell_buckets, (uint32_t) 8, pa, pb, part);
    b200::compute_launch(cu, {8, 128}, b200::array_to_buffer, pa, pb);
    b200::compute_launch(cu, {8, 128}, b200::retrieve_particle_buffer, part, prev, pb, next, pa, marks);
    (void) mv;
}
void common_calls(Ctx& cu, Partition<1>& part, Partition<1>& prev, GridBuffer& g0, GridBuffer& g1, ParticleArray& pa, int* marks, float* mv) {
    b200::compute_launch(cu, {8, 128}, b200::update_grid_velocity_query_max, (uint32_t) 8, g0, part, Duration(1e-4f), mv);
    b200::compute_launch(cu, {8, 64}, b200::clear_grid, g1);
    b200::compute_launch(cu, {8, 64}, b200::cell_bucket_to_block, (const int*) marks, (const int*) marks, marks, marks);
    b200::compute_launch(cu, {8, 128}, b200::compute_bin_capacity, (uint32_t) 8, (const int*) marks, marks);
    b200::compute_launch(cu, {8, 128}, b200::init_adv_bucket, (const int*) marks, marks);
    b200::compute_launch(cu, {8, 128}, b200::register_neighbor_blocks, (uint32_t) 8, part);
    b200::compute_launch(cu, {8, 128}, b200::register_exterior_blocks, (uint32_t) 8, part);
    b200::compute_launch(cu, {8, 128}, b200::mark_active_grid_blocks, (uint32_t) 8, (const GridBuffer) g1, marks);
    b200::compute_launch(cu, {8, 128}, b200::mark_active_particle_blocks, (uint32_t) 8, (const int*) marks, marks);
    b200::compute_launch(cu, {8, 128}, b200::exclusive_scan_inverse, 8, (const int*) marks, marks);
    b200::compute_launch(cu, {8, 128}, b200::update_partition, (uint32_t) 8, (const int*) marks, (const Partition<1>) prev, part);
    b200::compute_launch(cu, {8, 64}, b200::copy_selected_grid_blocks, (const ivec3*) part.active_keys, (const Partition<1>) part, (const int*) marks, g1, g0);
    b200::compute_launch(cu, {8, 128}, b200::activate_blocks, (uint32_t) 8, pa, part);
    b200::compute_launch(cu, {8, 128}, b200::rasterize, (uint32_t) 8, (const ParticleArray) pa, g0, (const Partition<1>) part, Duration(1e-4f), 1.f, std::array<float, 3> {0.f, 0.f, 0.f});
}
template void every_call<MaterialE::J_FLUID>(Ctx&, ParticleBuffer<MaterialE::J_FLUID>&, ParticleBuffer<MaterialE::J_FLUID>&, Partition<1>&, Partition<1>&, GridBuffer&, GridBuffer&, ParticleArray&, int*, float*);
template void every_call<MaterialE::FIXED_COROTATED>(Ctx&, ParticleBuffer<MaterialE::FIXED_COROTATED>&, ParticleBuffer<MaterialE::FIXED_COROTATED>&, Partition<1>&, Partition<1>&, GridBuffer&, GridBuffer&, ParticleArray&, int*, float*);
template void every_call<MaterialE::SAND>(Ctx&, ParticleBuffer<MaterialE::SAND>&, ParticleBuffer<MaterialE::SAND>&, Partition<1>&, Partition<1>&, GridBuffer&, GridBuffer&, ParticleArray&, int*, float*);
template void every_call<MaterialE::NACC>(Ctx&, ParticleBuffer<MaterialE::NACC>&, ParticleBuffer<MaterialE::NACC>&, Partition<1>&, Partition<1>&, GridBuffer&, GridBuffer&, ParticleArray&, int*, float*);
""")
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    inc = [f"-I{ref}/Library", f"-I{ref}/Projects/GMPM", f"-I{ref}/Externals/function_ref", f"-I{ref}/Externals/optional", f"-I{ref}/Externals/variant", f"-I{ROOT}/include"]
    cmd = [nvcc, "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "--expt-extended-lambda", "--expt-relaxed-constexpr", "-DQR_CUH", "-include", "chrono",
           *inc, "-c", "-o", str(tmp_path / "adapter_tu.o"), str(tu)]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-6000:]


def test_cmake_build_exports_the_same_abi(tmp_path):
    """CMakeLists.txt (sm_100a pinned) builds the same library for CMake consumers such as the reference tree."""
    import shutil
    import subprocess
    cmake = shutil.which("cmake")
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not cmake or not os.path.exists(nvcc):
        pytest.skip("cmake / nvcc not available")
    b = str(tmp_path / "build")
    r = subprocess.run([cmake, "-S", ROOT, "-B", b, f"-DCMAKE_CUDA_COMPILER={nvcc}"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "100a" in open(os.path.join(ROOT, "CMakeLists.txt")).read()
    r = subprocess.run([cmake, "--build", b, "-j", "4"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    lib = C.CDLL(os.path.join(b, "libclaymore_b200.so"))
    missing = [n for n in _declared_symbols() if not hasattr(lib, n)]
    assert not missing, missing
