"""ctypes binding of libwgbs_b200.so (include/wgbs_b200.h).

There is no CPU path: importing this module loads the CUDA library and raises if it is missing; creating a
``Context`` raises if no B200 is usable."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libwgbs_b200.so")


class PileupOpts(C.Structure):
    """mirror of wgbs_pileup_opts (include/wgbs_b200.h)"""
    _fields_ = [("min_cpg", C.c_int32), ("clip", C.c_int32), ("paired", C.c_int32), ("nanopore", C.c_int32),
                ("combine_mods", C.c_int32), ("np_thresh", C.c_float), ("cpc_call", C.c_char), ("keep_names", C.c_int32)]


class ViewOpts(C.Structure):
    """mirror of wgbs_view_opts (include/wgbs_b200.h)"""
    _fields_ = [("refid", C.c_int), ("min_mapq", C.c_int), ("exclude_flags", C.c_int), ("include_flags", C.c_int),
                ("beg", C.c_int64), ("end", C.c_int64), ("n_flag_eq", C.c_int), ("flag_eq", C.c_int * 4),
                ("read_group", C.c_char_p), ("iv_beg", C.c_void_p), ("iv_end", C.c_void_p), ("n_iv", C.c_size_t),
                ("iv_exclude", C.c_int), ("max_records", C.c_uint64), ("key_beg", C.c_int64), ("key_end", C.c_int64)]


class WgbsError(RuntimeError):
    """Any rc<0 from the C ABI (message from wgbs_last_error)."""


def _load():
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built (python -m wgbs_tools_b200.build). "
            "wgbs_tools_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, sz, u32, i32, u64 = C.c_void_p, C.c_size_t, C.c_uint32, C.c_int32, C.c_uint64
    sig = {
        "wgbs_abi_version": (C.c_int, []),
        "wgbs_last_error": (C.c_char_p, []),
        "wgbs_create": (vp, [C.c_int, vp]),
        "wgbs_destroy": (None, [vp]),
        "wgbs_sync": (C.c_int, [vp]),
        "wgbs_launch_count": (u64, [vp]),
        "wgbs_prof_enable": (C.c_int, [vp, C.c_int]),
        "wgbs_prof_report": (C.c_int, [vp, vp, sz]),
        "wgbs_dev_alloc": (C.c_int, [vp, sz, C.POINTER(vp)]),
        "wgbs_dev_free": (C.c_int, [vp, vp]),
        "wgbs_memcpy": (C.c_int, [vp, vp, vp, sz]),
        "wgbs_prefetch": (C.c_int, [vp, vp, vp, sz]),
        "wgbs_prefetch_wait": (C.c_int, [vp]),
        "wgbs_pats_from_text": (C.c_int, [vp, vp, sz, C.POINTER(vp)]),
        "wgbs_pats_count": (C.c_int, [vp, C.POINTER(u64), C.POINTER(u64)]),
        "wgbs_pats_download": (C.c_int, [vp, vp, vp, vp, vp, vp, vp]),
        "wgbs_pats_free": (None, [vp, vp]),
        "wgbs_pat2beta": (C.c_int, [vp, vp, u32, u32, vp, C.c_int]),
        "wgbs_trim": (C.c_int, [vp, vp, sz, C.c_int, vp]),
        "wgbs_pat2beta_text": (C.c_int, [vp, vp, sz, u32, u32, C.c_int, vp, vp]),
        "wgbs_homog": (C.c_int, [vp, vp, vp, vp, sz, vp, C.c_int, C.c_int, C.c_int, vp]),
        "wgbs_index_load": (C.c_int, [vp, vp, sz, u32, C.POINTER(vp)]),
        "wgbs_index_free": (None, [vp, vp]),
        "wgbs_pileup_sam": (C.c_int, [vp, vp, vp, sz, vp, C.POINTER(vp), vp]),
        "wgbs_pileup_sam_mbias": (C.c_int, [vp, vp, vp, sz, vp, C.POINTER(vp), vp, vp]),
        "wgbs_collapse": (C.c_int, [vp, vp]),
        "wgbs_collapse_long": (C.c_int, [vp, vp]),
        "wgbs_collapse_ex": (C.c_int, [vp, vp, C.c_int]),
        "wgbs_cview": (C.c_int, [vp, vp, vp, vp, sz, vp, vp, sz, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]),
        "wgbs_beta_to_blocks": (C.c_int, [vp, vp, C.c_int, sz, vp, vp, sz, C.c_int, vp, vp]),
        "wgbs_pats_format_long": (C.c_int, [vp, vp, C.c_char_p, vp, sz, C.POINTER(sz)]),
        "wgbs_pats_format": (C.c_int, [vp, vp, C.c_char_p, vp, sz, C.POINTER(sz)]),
        "wgbs_sort_pairs_u32": (C.c_int, [vp, vp, vp, sz]),
        "wgbs_segment": (C.c_int, [vp, vp, C.c_int, vp, sz, vp, C.c_int, C.c_int, u32, C.c_float, vp, vp]),
        "wgbs_glibc_log2_probe": (C.c_int, [vp, vp, sz, vp, vp]),
        "wgbs_bam_open": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(vp)]),
        "wgbs_bam_close": (None, [vp]),
        "wgbs_bam_nref": (C.c_int, [vp]),
        "wgbs_bam_ref_name": (C.c_char_p, [vp, C.c_int]),
        "wgbs_bam_header": (C.c_char_p, [vp]),
        "wgbs_bam_nrecords": (u64, [vp, C.c_int]),
        "wgbs_bam_view": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.POINTER(vp), C.POINTER(sz), C.POINTER(u64)]),
        "wgbs_bam_view_ex": (C.c_int, [vp, C.POINTER(ViewOpts), C.POINTER(vp), C.POINTER(sz), C.POINTER(u64)]),
        "wgbs_host_free": (None, [vp]),
        "wgbs_bam_open_part": (C.c_int, [vp, sz, C.c_int, vp, vp, C.c_int, u64, C.c_int, C.POINTER(vp), C.POINTER(u64)]),
        "wgbs_bam_last_record": (C.c_int, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int64)]),
        "wgbs_bam_first_key": (C.c_int, [vp, C.POINTER(ViewOpts), C.c_int, C.c_int64, C.POINTER(u64), C.POINTER(C.c_int)]),
        "wgbs_bam_inflated_bytes": (u64, [vp]),
        "wgbs_bam_probe": (C.c_int, [vp, sz, C.c_int, C.POINTER(u64), C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_int)]),
        "wgbs_bgzf_inflate": (C.c_int, [vp, vp, sz, C.POINTER(vp), C.POINTER(sz)]),
        "wgbs_dbam_open": (C.c_int, [vp, vp, sz, C.POINTER(vp)]),
        "wgbs_bgzf_index_build": (C.c_int, [vp, sz, C.POINTER(vp)]),
        "wgbs_bgzf_index_free": (None, [vp]),
        "wgbs_bgzf_index_blocks": (u64, [vp]),
        "wgbs_dbam_open_indexed": (C.c_int, [vp, vp, sz, vp, C.POINTER(vp)]),
        "wgbs_dbam_open_file": (C.c_int, [vp, C.c_char_p, C.POINTER(vp)]),
        "wgbs_dbam_close": (None, [vp, vp]),
        "wgbs_dbam_open_part": (C.c_int, [vp, vp, sz, C.c_int, vp, vp, C.c_int, u64, C.POINTER(vp), C.POINTER(u64)]),
        "wgbs_dbam_last_record": (C.c_int, [vp, vp, C.POINTER(C.c_int), C.POINTER(C.c_int64)]),
        "wgbs_dbam_first_key": (C.c_int, [vp, vp, C.POINTER(ViewOpts), C.c_int, C.c_int64, C.POINTER(u64), C.POINTER(C.c_int)]),
        "wgbs_dbam_nref": (C.c_int, [vp]),
        "wgbs_dbam_ref_name": (C.c_char_p, [vp, C.c_int]),
        "wgbs_dbam_header": (C.c_char_p, [vp]),
        "wgbs_dbam_nrecords": (u64, [vp, C.c_int]),
        "wgbs_dbam_inflated_bytes": (u64, [vp]),
        "wgbs_dbam_view": (C.c_int, [vp, vp, C.POINTER(ViewOpts), C.POINTER(vp), C.POINTER(sz), C.POINTER(u64)]),
        "wgbs_pileup_dbam": (C.c_int, [vp, vp, vp, C.POINTER(ViewOpts), vp, C.POINTER(vp), vp, vp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    return L, sig


lib, SIGNATURES = _load()
ABI_VERSION = 5          # WGBS_B200_ABI_VERSION of include/wgbs_b200.h these bindings mirror (struct layouts, signatures)
if lib.wgbs_abi_version() != ABI_VERSION:
    raise ImportError(f"{LIB_PATH} has ABI version {lib.wgbs_abi_version()}, these bindings expect {ABI_VERSION}: rebuild it "
                      "(python -m wgbs_tools_b200.build -f)")


def check(rc: int) -> None:
    if rc < 0:
        raise WgbsError(lib.wgbs_last_error().decode(errors="replace"))
