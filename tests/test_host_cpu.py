"""CPU-side checks of the product: the C ABI loads, exports every declared symbol, fails loudly
without a GPU, and its integer roll-over logic equals the oracle's (bit exact)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import juliaraytracingsw_b200 as swrt
from juliaraytracingsw_b200 import _lib, outputs
from oracle import outputs as oout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "swrt.h")).read()
    declared = set(re.findall(r"\b(swrt_[a-z_0-9]+)\s*\(", hdr))
    L = swrt.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert L.swrt_version() == 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(swrt.SwrtError):
        swrt.Problem(nx=64)


def test_bad_arguments_are_reported():
    L = swrt.lib()
    assert L.swrt_flow_create(None, None) == -1
    assert b"null" in L.swrt_last_error()
    d = _lib.FlowDesc(nx=100, ny=100, Lx=1.0, Ly=1.0, dt=0.1)
    h = C.c_void_p()
    assert L.swrt_flow_create(C.byref(d), C.byref(h)) == -3       # unsupported size, before touching CUDA
    assert b"powers of two" in L.swrt_last_error()


def test_sequenced_output_rollover_matches_oracle_K13():
    got = outputs.SequencedOutput("packets", 300)
    want = oout.SequencedOutput(lambda i: oout.packet_filename("packets", i), 300)
    oout.savepacketproblem(got); oout.savepacketproblem(want)
    for frame in range(0, 200):
        oout.write_packets(got, frame * 10, True)
        oout.write_packets(want, frame * 10, True)
    want_files = {k: v for k, v in want.files.items() if v}
    assert got.files == want_files
    assert got.files["packets.000001.jld2"][0] == "p/g/580"
    assert (got.file_index, got.current_writes) == (want.file_index, want.current_writes)


def test_collated_filename_K12():
    assert outputs.collated_filename("test_dir3/sin", 1) == "test_dir3/sin_00000001.out"


def test_twolayer_driver_lattice_and_parameters():
    """raytracing/TwoLayerRaytracing.jl:10-22 lattice (integer indexing bit-exact against the loop restatement) and
    simulation/Parameters.jl:6-24 derived parameters."""
    import numpy as np
    from oracle import raytrace as oray
    from juliaraytracingsw_b200 import twolayer
    for s in (1, 3, 20):
        xk, sign = oray.generate_initial_wavepackets_twolayer(2 * np.pi, np.sqrt(3.0), s)
        np.testing.assert_array_equal(twolayer.generate_initial_wavepackets(2 * np.pi, np.sqrt(3.0), s * s, s), xk)
        assert np.all(sign == 1)
    P = twolayer.Parameters()
    mu, b1, U = P.compute_parameters()
    assert abs(U - 0.1 / 7.5) < 1e-17 and abs(b1 - (4 / 225 + 1)) < 1e-15
    assert abs(mu - 2 * U * (0.36 / np.log(7.5 / 3.2)) * 15) < 1e-15
    assert abs(P.dt - 0.02 * (2 * np.pi / 512) / 0.1) < 1e-18


def test_julia_shim_is_consistent_with_the_header():
    """julia/SWRT.jl cannot be executed here (no Julia): check statically that every symbol it ccalls is declared in
    include/swrt.h and that its two descriptor structs list the header's fields in the header's order (as _lib.py does)."""
    hdr = open(os.path.join(ROOT, "include", "swrt.h")).read()
    jl = open(os.path.join(ROOT, "julia", "SWRT.jl")).read()
    declared = set(re.findall(r"\b(swrt_[a-z_0-9]+)\s*\(", hdr))
    called = set(re.findall(r"\(:(swrt_[a-z_0-9]+),\s*libswrt\)", jl))
    assert called and called <= declared, called - declared

    def header_fields(struct):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct, struct), hdr, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if decl:
                names += [n.strip() for n in decl.split(None, 1)[1].replace("long long", "").split(",")]
        return [n.split()[-1] for n in names]

    def julia_fields(struct):
        body = re.search(r"Base\.@kwdef struct %s\n(.*?)\nend" % struct, jl, re.S).group(1)
        body = re.sub(r"#.*", "", body)
        return re.findall(r"([A-Za-z_0-9]+)::C", body)

    for cname, jname, pycls in (("swrt_flow_desc", "FlowDesc", _lib.FlowDesc), ("swrt_packets_desc", "PacketsDesc", _lib.PacketsDesc)):
        want = header_fields(cname)
        assert julia_fields(jname) == want, (jname, julia_fields(jname), want)
        assert [f[0] for f in pycls._fields_] == want, (pycls, want)


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the oracle on the host cores) on a tiny grid: exactly one JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--nx", "64", "--sqrt-packets", "16"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    # under torchrun only rank 0 prints
    r2 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--nx", "64",
                         "--sqrt-packets", "16"], capture_output=True, text=True, timeout=300, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert r2.returncode == 0 and not [ln for ln in r2.stdout.splitlines() if ln.startswith("{")]


def test_header_is_plain_c():
    """include/swrt.h is the C ABI: it must compile as C99 (and as C++) on its own, with no CUDA or C++ types in the signatures."""
    import subprocess
    hdr = os.path.join(ROOT, "include", "swrt.h")
    for cmd in (["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", hdr],
                ["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", hdr]):
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    code = re.sub(r"/\*.*?\*/", "", open(hdr).read(), flags=re.S)            # declarations only
    for word in ("std::", "torch", "cudastream_t", "cudaevent_t", "double2", "at::", "template", "class "):   # no C++/CUDA/torch types
        assert word not in code.lower(), word


@pytest.mark.parametrize("max_writes", [8, 9, 10, 11, 12, 23])
@pytest.mark.parametrize("write_gradients", [False, True])
def test_frame_writer_is_accepted_by_the_reference_reader_K13_K14(tmp_path, max_writes, write_gradients):
    """f2 acceptance: the writer stores frames under the reference's keys (params/* then p/t, p/x, p/k, p/u[, p/g] per frame,
    raytracing/RaytracingDriver.jl:87-108), rolling to the next file after ANY key (utils/SequencedOutputs.jl:37-44,58-63), and
    the oracle's restatement of analysis/load_file.jl:89-160 puts the frames back together -- whichever key the roll-over fell on."""
    from juliaraytracingsw_b200.outputs import KeyedFile, SequencedOutput
    from oracle import outputs as oout
    rng = np.random.default_rng(max_writes)
    N, nframes = 13, 17
    out = SequencedOutput("packets", max_writes, store=True, directory=str(tmp_path))
    ref = oout.SequencedOutput(lambda i: oout.packet_filename("packets", i), max_writes)
    for key in ("f0", "Cg", "dt", "N", "k0", "ωsign"):
        out["params/" + key] = 1.0
    oout.savepacketproblem(ref)
    frames = []
    fr = -1
    # (the reference reader needs the LAST file to hold a p/t key whose frame is complete in that file -- there is no next file to
    # look into, analysis/load_file.jl:131-148 -- so the run is extended until the current file holds a whole frame)
    while fr + 1 < nframes or not any(k.startswith("p/t/") for k in ref.files[ref.current]):
        fr += 1
        step = 10 * fr
        t, x, k, u, g = 0.5 * fr, rng.standard_normal((N, 2)), rng.standard_normal((N, 2)), rng.standard_normal((N, 2)), rng.standard_normal((N, 4))
        frames.append((t, x, k, u))
        out[f"p/t/{step}"] = t
        out[f"p/x/{step}"] = x
        out[f"p/k/{step}"] = k
        out[f"p/u/{step}"] = u
        if write_gradients:
            out[f"p/g/{step}"] = g
        oout.write_packets(ref, step, write_gradients)
    out.close()
    assert out.files == {n: ks for n, ks in ref.files.items() if ks}          # same keys in the same files, in the same order
    nfiles = out.file_index + (1 if out.current_writes else 0)
    files = [KeyedFile.load(str(tmp_path / oout.packet_filename("packets", i))) for i in range(nfiles)]
    splits = sum(1 for f in files[:-1] if f"p/u/{f.keys('p/t')[-1]}" not in f)
    times, x, k, u = oout.load_packet_analysis_files_collated(lambda i: files[i], list(range(nfiles)), load_velocity=True)
    # the reader counts one frame per `p/t` key: all frames are there, in order, whichever key the roll-over fell on
    assert x.shape[0] == len(frames) >= nframes
    for fr, (t, xf, kf, uf) in enumerate(frames):
        np.testing.assert_array_equal(x[fr], xf)
        np.testing.assert_array_equal(k[fr], kf)
        np.testing.assert_array_equal(u[fr], uf)
    last = np.cumsum([len(f.keys("p/t")) for f in files]) - 1                   # (reader quirk: the last time of every file stays 0)
    np.testing.assert_array_equal(np.delete(times, last), np.delete(np.array([f[0] for f in frames]), last))
    if max_writes in ((8, 9, 12) if write_gradients else (8, 9, 10, 11)):
        assert splits > 0                                                       # frames did straddle files in these cases


def test_segment_index_by_multiply_high_is_exact():
    """csrc/passes.cuh `seg_magic`: the slab rows find the rank segment of column k as umulhi(k, ceil(2^32 / chunk)).  Exact for every
    chunk the library can produce (multiples of 16 up to 4096) and every column index below 8192."""
    import numpy as np
    k = np.arange(8192, dtype=np.uint64)
    for chunk in range(16, 4097, 16):
        magic = ((1 << 32) + chunk - 1) // chunk
        assert magic < (1 << 32)
        assert np.array_equal((k * np.uint64(magic)) >> np.uint64(32), k // np.uint64(chunk)), chunk
