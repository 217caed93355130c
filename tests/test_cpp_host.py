"""The C++ host API (include/feotensor.hpp) mirrors the reference's interface over the C ABI.

tests/cpp/gen_reference_port.py turns the reference's known-answer tests (tests/golden/reference_kat.json) into a
C++ program written against that API.  On CPU we check that it compiles and links against the in-tree library;
on a B200 (-m gpu) it runs: 96 cases x {double, float} + 60 integer cases x {i32, i64, u32} + bookkeeping and
panic preconditions, exact equality as in the reference's assert_eq!.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "build", "cpp", "test_reference_port.cpp")
EXE = os.path.join(ROOT, "build", "cpp", "test_reference_port")
LIBDIR = os.path.join(ROOT, "feotensor_b200", "lib")


def _build():
    gen = os.path.join(ROOT, "tests", "cpp", "gen_reference_port.py")
    subprocess.run([sys.executable, gen, SRC], check=True)
    deps = [SRC, os.path.join(ROOT, "include", "feotensor.hpp"), os.path.join(ROOT, "include", "feotensor_b200.h"),
            os.path.join(LIBDIR, "libfeotensor_b200.so")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"), SRC, "-L" + LIBDIR,
                        "-lfeotensor_b200", "-Wl,-rpath," + LIBDIR, "-o", EXE], check=True)
    return EXE


def test_cpp_port_compiles_and_links():
    exe = _build()
    assert os.access(exe, os.X_OK)


def test_cpp_peer_communicator_api_compiles(tmp_path):
    """The multi-GPU wrapper of the C++ host API (PeerWindow / PeerComm) instantiates against the C ABI for every dtype."""
    src = tmp_path / "peer_api.cpp"
    src.write_text("""
#include "feotensor.hpp"
template <class T> void use(feo::PeerComm &c, const feo::Tensor<T> &x, const feo::Shape &g) {
    (void)c.reduce(FEO_VAR, x, g, feo::Axes{0});
    (void)c.dot(x, x);
    (void)c.matmul_sharded_b(x, std::vector<const void *>{nullptr}, 4, 4);
}
int main(int argc, char **) {
    if (argc > 1000) {  // never runs: this program only has to compile and link
        auto ctx = std::make_shared<feo::Context>(0);
        auto win = std::make_shared<feo::PeerWindow>(ctx, 1 << 20);
        feo::PeerComm comm(ctx, 0, win, {win->export_handle()});
        comm.barrier();
        feo::Shape g({4, 4});
        use(comm, feo::Tensor<float>::zeros(g), g); use(comm, feo::Tensor<double>::zeros(g), g);
        use(comm, feo::Tensor<int32_t>::zeros(g), g); use(comm, feo::Tensor<int64_t>::zeros(g), g); use(comm, feo::Tensor<uint32_t>::zeros(g), g);
    }
    return 0;
}
""")
    exe = tmp_path / "peer_api"
    subprocess.run(["g++", "-std=c++17", "-O0", "-Wall", "-I" + os.path.join(ROOT, "include"), str(src), "-L" + LIBDIR, "-lfeotensor_b200",
                    "-Wl,-rpath," + LIBDIR, "-o", str(exe)], check=True)
    assert os.access(exe, os.X_OK)


def test_cpp_nccl_communicator_api_compiles(tmp_path):
    """The NCCL route of the C++ host API (feo::NcclComm over feo_nccl_*: NCCL issued inside the C ABI, no Python, no
    torch.distributed) instantiates against the C ABI for every dtype."""
    src = tmp_path / "nccl_api.cpp"
    src.write_text("""
#include "feotensor.hpp"
template <class T> void use(feo::NcclComm &c, const feo::Tensor<T> &x, const feo::Shape &g) {
    (void)c.reduce(FEO_VAR, x, g, feo::Axes{0});
    (void)c.reduce(FEO_MAX, x, g, feo::Axes{});
    (void)c.dot(x, x);
    (void)c.matmul_sharded_b(x, x, 4, 4);
}
int main(int argc, char **) {
    if (argc > 1000) {  // never runs: this program only has to compile and link
        auto ctx = std::make_shared<feo::Context>(0);
        const feo::NcclId id = feo::NcclComm::unique_id();   // rank 0; the host's own rendezvous carries it to the others
        feo::NcclComm comm(ctx, id, 0, 1);
        feo::Shape g({4, 4});
        use(comm, feo::Tensor<float>::zeros(g), g); use(comm, feo::Tensor<double>::zeros(g), g);
        use(comm, feo::Tensor<int32_t>::zeros(g), g); use(comm, feo::Tensor<int64_t>::zeros(g), g); use(comm, feo::Tensor<uint32_t>::zeros(g), g);
        return comm.rank() + comm.world();
    }
    return 0;
}
""")
    exe = tmp_path / "nccl_api"
    subprocess.run(["g++", "-std=c++17", "-O0", "-Wall", "-I" + os.path.join(ROOT, "include"), str(src), "-L" + LIBDIR, "-lfeotensor_b200",
                    "-Wl,-rpath," + LIBDIR, "-o", str(exe)], check=True)
    assert subprocess.run([str(exe)]).returncode == 0


@pytest.mark.gpu
def test_cpp_port_of_reference_tests_passes_on_gpu():
    exe = _build()
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    sys.stdout.write(out.stdout[-2000:])
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "0 failures" in out.stdout


def test_call_overhead_tool_compiles_and_links(tmp_path):
    """tools/call_overhead.cpp (per-call cost of the C ABI from a compiled host) builds against the in-tree library."""
    exe = tmp_path / "call_overhead"
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tools", "call_overhead.cpp"),
                    "-L" + LIBDIR, "-lfeotensor_b200", "-Wl,-rpath," + LIBDIR, "-o", str(exe)], check=True)
    assert os.access(exe, os.X_OK)
