"""CPU: librrl.so builds (nvcc cross-compiles sm_100a without a GPU), loads, and exports every entry point that
include/rrl.h declares; host-only entry points (arena layout, CPython-compatible seeding) behave; compute
entry points refuse to run without a device instead of falling back."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_declared_symbol_is_exported(native):
    hdr = open(os.path.join(ROOT, "include", "rrl.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(rrl_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 25
    lib = ctypes.CDLL(native.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert native.version() == 6


def test_arena_layout_matches_reference_parameter_shapes(native):
    cfg = native.agent_config(max_batch=256)
    shapes = {
        "critic": [(256, 4), (256,), (256, 256), (256,), (1, 256), (1,)] * 2,
        "policy": [(256, 2), (256,), (256, 256), (256,), (2, 256), (2,), (2, 256), (2,)],
        "qrisk": [(4,), (4,)] + [(256, 4), (256,), (256, 256), (256,), (1, 256), (1,)] * 2,
        "recovery": [(2,), (256, 2), (256,), (256, 256), (256,), (2, 256), (2,)],
    }
    shapes["critic_target"] = shapes["critic"]
    shapes["qrisk_target"] = shapes["qrisk"]
    seen = []
    for net, want in shapes.items():
        idx = native.NET_NAMES.index(net)
        assert native.agent_num_tensors(idx) == len(want)
        for i, shp in enumerate(want):
            off, rows, cols = native.agent_tensor_info(cfg, idx, i)
            assert ((rows, cols) if cols else (rows,)) == shp
            assert off % 4 == 0
            seen.append((off, int(np.prod(shp))))
    seen.sort()
    for (o0, n0), (o1, _) in zip(seen, seen[1:]):
        assert o0 + n0 <= o1                                   # no overlap
    assert native.agent_arena_floats(cfg) > seen[-1][0]
    with pytest.raises(native.RRLError):
        native.agent_tensor_info(cfg, 0, 99)
    bad = native.agent_config(hidden=128)
    assert native.agent_arena_floats(bad) < 0                  # kernels are specialised for hidden 256


def test_seed_matches_cpython(native):
    import random
    from oracle.replay import MT19937
    for seed in (0, 1, 123456, 2 ** 40 + 5, -3):
        st = native.mt19937_seed(seed).numpy().view(np.uint32)
        assert np.array_equal(st, MT19937(seed).state625())
        random.seed(seed)
        ref = random.getstate()[1]
        assert list(st[:624]) == list(ref[:624])


def test_no_cpu_fallback(native):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(native.RRLError):
        native.require_cuda()
    from recovery_rl.engine import VecEngine
    with pytest.raises(native.RRLError):
        VecEngine("maze", 128)
    cpu = torch.zeros(8)
    with pytest.raises(native.RRLError):
        native.p(cpu, "f32")


def test_peer_table_packing_and_argument_checks(native):
    """rrl_peers_t (the peer-memory gradient sum of the sharded run): packing of the host-side table and the argument
    validation of the entry points, which happens before anything is launched (no GPU needed)."""
    P = native.make_peers(1, [0x1000, 0x2000, 0x3000], [0x10, 0x20, 0x30])
    assert (P.world, P.rank) == (3, 1)
    assert list(P.arena)[:4] == [0x1000, 0x2000, 0x3000, 0] and list(P.signal)[:4] == [0x10, 0x20, 0x30, 0]
    assert ctypes.sizeof(P) == 8 + 8 * 8 + 8 * 8 + 8 + 8             # int32 world, rank; uint64 arena[8], signal[8], epoch, mc_arena
    assert native.make_peers(0, [1, 2], [3, 4], epoch_ptr=0x40).epoch == 0x40 and P.epoch == 0
    with pytest.raises(native.RRLError):
        native.make_peers(0, list(range(9)), list(range(9)))           # one node: at most 8 ranks
    with pytest.raises(native.RRLError):
        native.make_peers(0, [1, 2], [1])
    lib = native.lib()
    assert lib.rrl_peer_barrier(None, None, None, None) < 0
    assert b"bad argument" in lib.rrl_last_error()
    bad = native.make_peers(0, [0x1000], [0x10])
    bad.rank = 5                                                        # rank outside [0, world)
    one = (ctypes.c_int64 * 1)()
    assert lib.rrl_peer_barrier(ctypes.byref(bad), one, one, None) < 0
    cfg = native.agent_config(max_batch=256)
    assert lib.rrl_sac_apply_p2p(ctypes.byref(cfg), None, None, None, None) < 0


def test_entry_points_reject_bad_arguments_before_launching(native):
    """error behaviour of the boundary: every call returns a negative code and leaves a message in rrl_last_error()
    instead of launching on bad input (checked here without a device)."""
    C = ctypes
    lib = native.lib()
    cfg = native.agent_config(max_batch=256)
    cases = [
        ("rrl_sac_apply", (None, None, None, None), b"null"),
        ("rrl_qrisk_apply", (C.byref(cfg), None, None, None), b"null"),
        ("rrl_recovery_apply", (C.byref(cfg), None, None, None), b"null"),
        ("rrl_agent_refresh", (None, None, None), b"null"),
        ("rrl_hard_update", (C.byref(cfg), None, C.c_int(9), C.c_int(9), None), b"null"),
        ("rrl_replay_flag_count", (None, C.c_int64(0), C.c_int32(512), None, None), b"null"),
        ("rrl_dyn_train_sync", (None, None, None), b"null"),
        ("rrl_mt19937_seed_host", (None, C.c_int(0), None), b"bad key"),
        ("rrl_counters_advance", (None, C.c_int64(1), C.c_int64(1), C.c_int64(1), C.c_int(1), C.c_int(1), None), b"null"),
    ]
    for name, args, text in cases:
        assert getattr(lib, name)(*args) < 0, name
        assert text in lib.rrl_last_error(), (name, lib.rrl_last_error())
    bad = native.agent_config(max_batch=256)
    bad.max_batch = 7
    lib.rrl_agent_arena_floats.restype = C.c_int64
    assert lib.rrl_agent_arena_floats(C.byref(bad)) < 0 and b"max_batch" in lib.rrl_last_error()
    assert lib.rrl_agent_num_tensors(C.c_int(99)) < 0


def test_header_is_plain_c_and_links_against_the_library(native, tmp_path):
    """include/rrl.h compiles as C99 (no C++ or torch types in the boundary) and a C program that takes the address of every
    declared entry point links against librrl.so."""
    import re
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = os.path.join(root, "include", "rrl.h")
    text = open(hdr).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = sorted(set(re.findall(r"\b(rrl_[a-z0-9_]+)\s*\(", text)))
    assert len(names) > 40
    src = tmp_path / "link_all.c"
    src.write_text('#include "rrl.h"\n#include <stdio.h>\nint main(void) {\n    const void* f[] = {%s};\n'
                   '    printf("%%d %%d\\n", (int)(sizeof(f) / sizeof(f[0])), rrl_version());\n    return 0;\n}\n'
                   % ", ".join("(const void*)%s" % n for n in names))
    lib_dir = os.path.join(root, "recovery-rl_b200")
    exe = tmp_path / "link_all"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic-errors", "-Wno-pedantic", "-I", os.path.join(root, "include"),
                           str(src), "-o", str(exe), "-L", lib_dir, "-l:librrl.so", "-Wl,-rpath," + lib_dir])
    out = subprocess.check_output([str(exe)], text=True).split()
    assert int(out[0]) == len(names) and int(out[1]) == native.version()
