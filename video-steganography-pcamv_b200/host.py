"""ctypes mirror of include/pcamv.h.

Python stands in for the reference's C host here only because the parity tests and bench.py are
Python; the reference encoder binds the very same C symbols.  Nothing in this file
computes anything: every method is one C-ABI call.  If the CUDA library is missing or no GPU is
present the calls raise — there is no CPU path.
"""
import ctypes as C
import os

import numpy as np

from . import build

MAX_MVC = 10


class PcamvError(RuntimeError):
    pass


class Cfg(C.Structure):
    _fields_ = [("abi_version", C.c_int), ("device", C.c_int), ("width", C.c_int), ("height", C.c_int),
                ("me_method", C.c_int), ("me_range", C.c_int), ("subpel_refine", C.c_int), ("chroma_me", C.c_int),
                ("max_refs", C.c_int), ("mv_range", C.c_int), ("b_cabac", C.c_int), ("b_fast_pskip", C.c_int),
                ("b_dct_decimate", C.c_int), ("analyse_inter", C.c_int), ("chroma_qp_offset", C.c_int),
                ("rows_per_cta", C.c_int), ("pass2_elide", C.c_int), ("no_deblock", C.c_int), ("deblock_alpha_c0_offset", C.c_int),
                ("deblock_beta_offset", C.c_int), ("reserved", C.c_int * 3)]


class QpTables(C.Structure):
    _fields_ = [("qp", C.c_int), ("lambda_", C.c_int), ("lambda2_chroma", C.c_int), ("chroma_qp", C.c_int),
                ("cost_mv", C.c_void_p), ("cost_ref", C.c_void_p),
                ("quant4_mf", C.c_void_p * 2), ("quant4_bias", C.c_void_p * 2), ("dequant4_mf", C.c_void_p * 2)]


# numpy views of the POD records (layout checked against sizeof in load_library)
ME_CALL_DTYPE = np.dtype([
    ("mode", "<i4"), ("mb_x", "<i4"), ("mb_y", "<i4"), ("xoff", "<i4"), ("yoff", "<i4"), ("i_pixel", "<i4"),
    ("ref_slot", "<i4"), ("i_ref_cost", "<i4"),
    ("mv_min_fpel", "<i4", 2), ("mv_max_fpel", "<i4", 2), ("mv_min_spel", "<i4", 2), ("mv_max_spel", "<i4", 2),
    ("i_mvc", "<i4"), ("has_thresh", "<i4"), ("thresh_in", "<i4"),
    ("mvp", "<i2", 2), ("mvc", "<i2", (MAX_MVC, 2)), ("mv_in", "<i2", 2),
    ("cost_in", "<i4"), ("cost_mv_in", "<i4")], align=True)
ME_RESULT_DTYPE = np.dtype([("mv", "<i2", 2), ("cost", "<i4"), ("cost_mv", "<i4"), ("thresh_out", "<i4")], align=True)

LOG_MAX = 112
LOG_ENTRY_DTYPE = np.dtype([("kind", "i1"), ("i_pixel", "i1"), ("i_ref", "i1"), ("pad", "i1"), ("mv", "<i2", 2),
                            ("cost", "<i4"), ("cost_mv", "<i4")], align=True)
RECON_PATCH_DTYPE = np.dtype([("mb_xy", "<i4"), ("nnz", "<u2"), ("pad", "<u2"), ("y", "u1", 256), ("u", "u1", 64), ("v", "u1", 64)])
MB_OUT_DTYPE = np.dtype([("type", "i1"), ("partition", "i1"), ("n_part", "i1"), ("early_skip", "i1"), ("ref", "i1", 4),
                         ("mv", "<i2", (16, 2)),
                         ("part", [("mv", "<i2", 2), ("mvp", "<i2", 2), ("ref", "i1"), ("i_pixel", "i1"), ("xoff", "i1"),
                                   ("yoff", "i1")], 4),
                         ("n_log", "<i4"), ("pskip_mv", "<i2", 2)], align=True)
PASS1_MB_DTYPE = np.dtype([("type", "<i4"), ("partition", "<i4"), ("used", "u1"), ("sub", "u1", 4), ("ref", "i1", 16),
                           ("mv", "<i2", (16, 2)), ("mv_stego", "<i2", (16, 2))], align=True)


class FrameIn(C.Structure):
    _fields_ = [("pass_", C.c_int32), ("n_ref", C.c_int32), ("ref_slot", C.c_int32 * 16), ("ref_poc", C.c_int32 * 16),
                ("cur_poc", C.c_int32), ("col_n_ref", C.c_int32), ("col_inv_ref_poc", C.c_int32 * 16),
                ("col_ref8", C.c_void_p), ("col_mv4", C.c_void_p), ("pass1", C.c_void_p), ("filp", C.c_void_p),
                ("n_filp", C.c_int32), ("cost_table", C.c_int32), ("stale_mv", (C.c_int16 * 2) * 16), ("device_forced", C.c_int32)]


EXPORTS = ["pcamv_open", "pcamv_close", "pcamv_last_error", "pcamv_abi_version", "pcamv_set_qp_tables",
           "pcamv_put_fenc", "pcamv_put_ref", "pcamv_put_ref_planes", "pcamv_get_ref_plane", "pcamv_plane_bytes",
           "pcamv_plane_stride", "pcamv_me_search_batch", "pcamv_me_batch_upload", "pcamv_me_batch_run",
           "pcamv_me_batch_download", "pcamv_launch_count", "pcamv_int_peak",
           "pcamv_analyse_p", "pcamv_frame_upload", "pcamv_frame_run", "pcamv_frame_download", "pcamv_frame_trace", "pcamv_log_stride",
           "pcamv_analyse_p_batch", "pcamv_frame_run_batch", "pcamv_host_alloc", "pcamv_host_free", "pcamv_set_pass2_elide", "pcamv_set_conformant",
           "pcamv_group_create", "pcamv_group_destroy", "pcamv_group_analyse_p", "pcamv_group_leave", "pcamv_stc_embed",
           "pcamv_embed_prepare", "pcamv_embed_stc", "pcamv_embed_download", "pcamv_reconstruct_ref",
           "pcamv_analyse_p_begin", "pcamv_analyse_p_batch_begin", "pcamv_group_analyse_p_begin", "pcamv_analyse_p_rows"]

_lib = None


def load_library(path=None):
    """dlopen libpcamv_cuda.so (built in-tree by build.build_cuda) and declare the prototypes."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or os.environ.get("PCAMV_LIB") or build.LIB        # PCAMV_LIB: try an experimental build of the library
    if not os.path.exists(path):
        raise PcamvError("libpcamv_cuda.so is not built (%s); run __graft_entry__.build()" % path)
    lib = C.CDLL(path)
    vp, ip = C.c_void_p, C.c_int
    lib.pcamv_open.argtypes = [C.POINTER(vp), C.POINTER(Cfg)]; lib.pcamv_open.restype = ip
    lib.pcamv_close.argtypes = [vp]; lib.pcamv_close.restype = None
    lib.pcamv_last_error.argtypes = [vp]; lib.pcamv_last_error.restype = C.c_char_p
    lib.pcamv_abi_version.argtypes = []; lib.pcamv_abi_version.restype = ip
    lib.pcamv_set_qp_tables.argtypes = [vp, C.POINTER(QpTables)]; lib.pcamv_set_qp_tables.restype = ip
    lib.pcamv_put_fenc.argtypes = [vp, vp, vp, vp, ip, ip]; lib.pcamv_put_fenc.restype = ip
    lib.pcamv_put_ref.argtypes = [vp, ip, ip, vp, vp, vp, ip, ip]; lib.pcamv_put_ref.restype = ip
    lib.pcamv_put_ref_planes.argtypes = [vp, ip, ip, C.POINTER(vp), vp, vp]; lib.pcamv_put_ref_planes.restype = ip
    lib.pcamv_get_ref_plane.argtypes = [vp, ip, ip, vp]; lib.pcamv_get_ref_plane.restype = ip
    lib.pcamv_plane_bytes.argtypes = [vp, ip]; lib.pcamv_plane_bytes.restype = C.c_size_t
    lib.pcamv_plane_stride.argtypes = [vp, ip]; lib.pcamv_plane_stride.restype = ip
    lib.pcamv_me_search_batch.argtypes = [vp, vp, ip, vp]; lib.pcamv_me_search_batch.restype = ip
    lib.pcamv_me_batch_upload.argtypes = [vp, vp, ip]; lib.pcamv_me_batch_upload.restype = ip
    lib.pcamv_me_batch_run.argtypes = [vp, ip, C.POINTER(C.c_float)]; lib.pcamv_me_batch_run.restype = ip
    lib.pcamv_me_batch_download.argtypes = [vp, vp, ip]; lib.pcamv_me_batch_download.restype = ip
    lib.pcamv_launch_count.argtypes = [vp]; lib.pcamv_launch_count.restype = C.c_longlong
    lib.pcamv_int_peak.argtypes = [vp, C.POINTER(C.c_double)]; lib.pcamv_int_peak.restype = ip
    lib.pcamv_analyse_p.argtypes = [vp, C.POINTER(FrameIn), vp, vp]; lib.pcamv_analyse_p.restype = ip
    lib.pcamv_frame_upload.argtypes = [vp, C.POINTER(FrameIn)]; lib.pcamv_frame_upload.restype = ip
    lib.pcamv_analyse_p_begin.argtypes = [vp, C.POINTER(FrameIn), vp, vp]; lib.pcamv_analyse_p_begin.restype = ip
    lib.pcamv_analyse_p_batch_begin.argtypes = [C.POINTER(vp), C.POINTER(C.POINTER(FrameIn)), ip, C.POINTER(vp), C.POINTER(vp)]
    lib.pcamv_analyse_p_batch_begin.restype = ip
    lib.pcamv_group_analyse_p_begin.argtypes = [vp, vp, C.POINTER(FrameIn), vp, vp]; lib.pcamv_group_analyse_p_begin.restype = ip
    lib.pcamv_analyse_p_rows.argtypes = [vp, ip, C.POINTER(ip)]; lib.pcamv_analyse_p_rows.restype = ip
    lib.pcamv_frame_run.argtypes = [vp, ip, ip, C.POINTER(C.c_float), C.POINTER(C.c_float)]; lib.pcamv_frame_run.restype = ip
    lib.pcamv_frame_download.argtypes = [vp, vp, vp]; lib.pcamv_frame_download.restype = ip
    lib.pcamv_frame_trace.argtypes = [vp, ip, vp]; lib.pcamv_frame_trace.restype = ip
    lib.pcamv_log_stride.argtypes = [vp]; lib.pcamv_log_stride.restype = ip
    lib.pcamv_set_pass2_elide.argtypes = [vp, ip]; lib.pcamv_set_pass2_elide.restype = ip
    lib.pcamv_set_conformant.argtypes = [vp, ip]; lib.pcamv_set_conformant.restype = ip
    lib.pcamv_group_create.argtypes = [C.POINTER(vp), ip]; lib.pcamv_group_create.restype = ip
    lib.pcamv_group_destroy.argtypes = [vp]; lib.pcamv_group_destroy.restype = None
    lib.pcamv_group_analyse_p.argtypes = [vp, vp, C.POINTER(FrameIn), vp, vp]; lib.pcamv_group_analyse_p.restype = ip
    lib.pcamv_group_leave.argtypes = [vp]; lib.pcamv_group_leave.restype = ip
    lib.pcamv_stc_embed.argtypes = [vp, vp, ip, vp, ip, vp, vp, ip, vp, ip, vp, ip]; lib.pcamv_stc_embed.restype = ip
    lib.pcamv_embed_prepare.argtypes = [vp, C.POINTER(C.c_int)]; lib.pcamv_embed_prepare.restype = ip
    lib.pcamv_embed_stc.argtypes = [vp, vp, ip, ip, vp, ip, vp, ip, C.c_double, vp]; lib.pcamv_embed_stc.restype = ip
    lib.pcamv_embed_download.argtypes = [vp, vp, vp, vp, vp, vp]; lib.pcamv_embed_download.restype = ip
    lib.pcamv_reconstruct_ref.argtypes = [vp, ip, ip, ip, vp, ip]; lib.pcamv_reconstruct_ref.restype = ip
    lib.pcamv_host_alloc.argtypes = [C.c_size_t]; lib.pcamv_host_alloc.restype = vp
    lib.pcamv_host_free.argtypes = [vp]; lib.pcamv_host_free.restype = None
    lib.pcamv_analyse_p_batch.argtypes = [C.POINTER(vp), C.POINTER(C.POINTER(FrameIn)), ip, C.POINTER(vp), C.POINTER(vp)]
    lib.pcamv_analyse_p_batch.restype = ip
    lib.pcamv_frame_run_batch.argtypes = [C.POINTER(vp), ip, ip, ip, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.pcamv_frame_run_batch.restype = ip
    if path == build.LIB:
        _lib = lib
    return lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


_pinned_keep = []


def pinned_array(shape, dtype):
    """numpy array over page-locked memory from pcamv_host_alloc (kept alive for the life of the process)."""
    lib = load_library()
    dt = np.dtype(dtype)
    n = int(np.prod(shape)) * dt.itemsize
    p = lib.pcamv_host_alloc(n)
    if not p:
        raise PcamvError("pcamv_host_alloc(%d) failed" % n)
    buf = (C.c_uint8 * n).from_address(p)
    _pinned_keep.append(buf)
    return np.frombuffer(buf, dtype=dt).reshape(shape)


def pinned_copy(a):
    out = pinned_array(a.shape, a.dtype)
    out[...] = a
    return out


class PcamvContext:
    """One encoder's GPU context (mirrors one x264_t)."""

    def __init__(self, width, height, me_method=1, me_range=16, subpel_refine=5, chroma_me=1, max_refs=1,
                 mv_range=512, b_cabac=1, b_fast_pskip=1, b_dct_decimate=1, analyse_inter=0x113, device=0, rows_per_cta=1, pass2_elide=0):
        self.lib = load_library()
        cfg = Cfg()
        cfg.abi_version = self.lib.pcamv_abi_version()
        cfg.device = device
        cfg.width, cfg.height = width, height
        cfg.me_method, cfg.me_range, cfg.subpel_refine, cfg.chroma_me = me_method, me_range, subpel_refine, chroma_me
        cfg.max_refs, cfg.mv_range, cfg.b_cabac, cfg.b_fast_pskip = max_refs, mv_range, b_cabac, b_fast_pskip
        cfg.b_dct_decimate, cfg.analyse_inter = b_dct_decimate, analyse_inter
        cfg.rows_per_cta = rows_per_cta
        cfg.pass2_elide = pass2_elide
        self.cfg = cfg
        self.handle = C.c_void_p()
        if self.lib.pcamv_open(C.byref(self.handle), C.byref(cfg)) != 0:
            raise PcamvError(self.lib.pcamv_last_error(None).decode())
        self.width, self.height = width, height
        self.log_stride = int(self.lib.pcamv_log_stride(self.handle))
        self._keep = []

    def close(self):
        if self.handle:
            self.lib.pcamv_close(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise PcamvError(self.lib.pcamv_last_error(self.handle).decode())

    # -- tables ---------------------------------------------------------------------------------
    def set_qp_tables(self, qp, lam, cost_mv, cost_ref=None, lambda2_chroma=0, chroma_qp=0,
                      quant4_mf=(None, None), quant4_bias=(None, None), dequant4_mf=(None, None)):
        t = QpTables()
        t.qp, t.lambda_, t.lambda2_chroma, t.chroma_qp = qp, lam, lambda2_chroma, chroma_qp
        keep = []

        def arr(a, dt, n):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=dt)
            assert a.size == n, (a.size, n)
            keep.append(a)
            return a.ctypes.data

        t.cost_mv = arr(cost_mv, np.int16, 32769)
        t.cost_ref = arr(cost_ref, np.uint16, 99)
        for i in range(2):
            t.quant4_mf[i] = arr(quant4_mf[i], np.uint16, 16)
            t.quant4_bias[i] = arr(quant4_bias[i], np.uint16, 16)
            t.dequant4_mf[i] = arr(dequant4_mf[i], np.int32, 96)
        self._check(self.lib.pcamv_set_qp_tables(self.handle, C.byref(t)))

    # -- frames ---------------------------------------------------------------------------------
    def put_fenc(self, y, u, v):
        y, u, v = (np.ascontiguousarray(p, dtype=np.uint8) for p in (y, u, v))
        self._check(self.lib.pcamv_put_fenc(self.handle, _ptr(y), _ptr(u), _ptr(v), y.strides[0], u.strides[0]))

    def put_ref(self, slot, poc, y, u, v):
        y, u, v = (np.ascontiguousarray(p, dtype=np.uint8) for p in (y, u, v))
        self._check(self.lib.pcamv_put_ref(self.handle, slot, poc, _ptr(y), _ptr(u), _ptr(v), y.strides[0], u.strides[0]))

    def put_ref_planes(self, slot, poc, luma4, u, v):
        luma4 = [np.ascontiguousarray(p, dtype=np.uint8) for p in luma4]
        u, v = np.ascontiguousarray(u, dtype=np.uint8), np.ascontiguousarray(v, dtype=np.uint8)
        for p in luma4:
            assert p.nbytes == self.plane_bytes(0), (p.nbytes, self.plane_bytes(0))
        assert u.nbytes == self.plane_bytes(4) and v.nbytes == self.plane_bytes(4)
        ptrs = (C.c_void_p * 4)(*[p.ctypes.data for p in luma4])
        self._check(self.lib.pcamv_put_ref_planes(self.handle, slot, poc, ptrs, _ptr(u), _ptr(v)))

    def plane_bytes(self, plane):
        return int(self.lib.pcamv_plane_bytes(self.handle, plane))

    def plane_stride(self, plane):
        return int(self.lib.pcamv_plane_stride(self.handle, plane))

    def get_ref_plane(self, slot, plane):
        out = np.empty(self.plane_bytes(plane), dtype=np.uint8)
        self._check(self.lib.pcamv_get_ref_plane(self.handle, slot, plane, _ptr(out)))
        return out.reshape(-1, self.plane_stride(plane))

    def stc_embed(self, cover, message, rho, cols_short, cols_long, height=10):
        """pcamv_stc_embed: returns the stego bit vector, or None when the message is not embeddable in this cover."""
        cover = np.ascontiguousarray(cover, dtype=np.uint8); message = np.ascontiguousarray(message, dtype=np.uint8)
        rho = np.ascontiguousarray(rho, dtype=np.float32)
        cs = np.ascontiguousarray(cols_short, dtype=np.uint32); cl = np.ascontiguousarray(cols_long, dtype=np.uint32)
        stego = np.zeros(len(cover), dtype=np.uint8)
        rc = self.lib.pcamv_stc_embed(self.handle, _ptr(cover), len(cover), _ptr(message), len(message), _ptr(rho), _ptr(stego),
                                      height, _ptr(cs), len(cs), _ptr(cl), len(cl))
        if rc < 0:
            self._check(rc)
        return stego if rc == 0 else None

    # -- reference frame built on the device ---------------------------------------------------------
    def reconstruct_ref(self, slot, poc, pass_, patches=None):
        """pcamv_reconstruct_ref: the frame of the last analysed (final) pass into reference slot `slot`."""
        if patches is not None and len(patches):
            patches = np.ascontiguousarray(patches, dtype=RECON_PATCH_DTYPE)
            self._check(self.lib.pcamv_reconstruct_ref(self.handle, slot, poc, pass_, _ptr(patches), len(patches)))
        else:
            self._check(self.lib.pcamv_reconstruct_ref(self.handle, slot, poc, pass_, None, 0))

    # -- embed stage on the device -----------------------------------------------------------------
    def embed_prepare(self):
        """pcamv_embed_prepare: cover / rho assembly from the pass-1 results in HBM; returns the cover length."""
        n = C.c_int()
        self._check(self.lib.pcamv_embed_prepare(self.handle, C.byref(n)))
        self._emb_len = int(n.value)
        return self._emb_len

    def embed_stc(self, message, cols_short, cols_long, height=10, total=-1.0):
        """pcamv_embed_stc on the device-resident cover / rho; returns (rc, stego) with rc 0 = embedded, 1 = not embeddable."""
        stego = np.zeros(max(self._emb_len, 1), dtype=np.uint8)
        if message is None or len(message) == 0:
            rc = self.lib.pcamv_embed_stc(self.handle, None, 0, height, None, 0, None, 0, C.c_double(total), _ptr(stego))
        else:
            message = np.ascontiguousarray(message, dtype=np.uint8)
            cs = np.ascontiguousarray(cols_short, dtype=np.uint32); cl = np.ascontiguousarray(cols_long, dtype=np.uint32)
            rc = self.lib.pcamv_embed_stc(self.handle, _ptr(message), len(message), height, _ptr(cs), len(cs), _ptr(cl), len(cl),
                                          C.c_double(total), _ptr(stego))
        if rc < 0:
            self._check(rc)
        return rc, stego[:self._emb_len]

    def embed_download(self, want_stego=True):
        """cover, rho, stego, filp (length entries each) and the per-macroblock info.cache[] records."""
        n = max(self._emb_len, 1)
        n_mb = (self.width // 16) * (self.height // 16)
        cover = np.zeros(n, np.uint8); rho = np.zeros(n, np.float32); stego = np.zeros(n, np.uint8); filp = np.zeros(n, np.int8)
        p1 = np.zeros(n_mb, dtype=PASS1_MB_DTYPE)
        self._check(self.lib.pcamv_embed_download(self.handle, _ptr(cover), _ptr(rho), _ptr(stego) if want_stego else None,
                                                  _ptr(filp) if want_stego else None, _ptr(p1)))
        k = self._emb_len
        return cover[:k], rho[:k], stego[:k], filp[:k], p1

    def get_integral(self, slot):
        """The integral plane of a slot (--me esa / tesa contexts): uint16 [rows, stride_y]."""
        out = np.empty(self.plane_bytes(6), dtype=np.uint8)
        self._check(self.lib.pcamv_get_ref_plane(self.handle, slot, 6, _ptr(out)))
        return out.view(np.uint16).reshape(-1, self.plane_stride(0))

    # -- search seam ------------------------------------------------------------------------------
    def me_search_batch(self, calls):
        calls = np.ascontiguousarray(calls, dtype=ME_CALL_DTYPE)
        res = np.zeros(len(calls), dtype=ME_RESULT_DTYPE)
        self._check(self.lib.pcamv_me_search_batch(self.handle, _ptr(calls), len(calls), _ptr(res)))
        return res

    def me_batch_upload(self, calls):
        calls = np.ascontiguousarray(calls, dtype=ME_CALL_DTYPE)
        self._check(self.lib.pcamv_me_batch_upload(self.handle, _ptr(calls), len(calls)))
        self._batch_n = len(calls)

    def me_batch_run(self, iters=1):
        ms = C.c_float()
        self._check(self.lib.pcamv_me_batch_run(self.handle, iters, C.byref(ms)))
        return float(ms.value)

    def me_batch_download(self):
        res = np.zeros(self._batch_n, dtype=ME_RESULT_DTYPE)
        self._check(self.lib.pcamv_me_batch_download(self.handle, _ptr(res), self._batch_n))
        return res

    # -- frame seam --------------------------------------------------------------------------------
    def _frame_in(self, pass_, ref_slots, ref_pocs, cur_poc, col_n_ref=0, col_inv_ref_poc=None, col_ref8=None,
                  col_mv4=None, pass1=None, filp=None, cost_table=True, stale_mv=None, device_forced=False):
        fi = FrameIn()
        keep = []
        fi.pass_, fi.n_ref, fi.cur_poc, fi.col_n_ref = pass_, len(ref_slots), cur_poc, col_n_ref
        for i, (s, p) in enumerate(zip(ref_slots, ref_pocs)):
            fi.ref_slot[i], fi.ref_poc[i] = int(s), int(p)
        for i, v in enumerate(col_inv_ref_poc if col_inv_ref_poc is not None else []):
            fi.col_inv_ref_poc[i] = int(v)
        if col_n_ref > 0:
            a = np.ascontiguousarray(col_ref8, dtype=np.int8); b = np.ascontiguousarray(col_mv4, dtype=np.int16)
            n_mb = (self.width // 16) * (self.height // 16)
            assert a.size == 4 * n_mb and b.size == 32 * n_mb, (a.size, b.size, n_mb)
            keep += [a, b]
            fi.col_ref8, fi.col_mv4 = a.ctypes.data, b.ctypes.data
        if pass1 is not None:
            a = np.ascontiguousarray(pass1, dtype=PASS1_MB_DTYPE)
            f = np.ascontiguousarray(filp if filp is not None else [], dtype=np.int8)
            keep += [a, f]
            fi.pass1, fi.filp, fi.n_filp = a.ctypes.data, f.ctypes.data, len(f)
        fi.cost_table = int(bool(cost_table))
        fi.device_forced = int(bool(device_forced))
        if stale_mv is not None:
            for i in range(16):
                fi.stale_mv[i][0], fi.stale_mv[i][1] = int(stale_mv[i][0]), int(stale_mv[i][1])
        return fi, keep

    def alloc_outputs(self, pinned=False):
        """(mb records [n_mb], log [n_mb, log_stride]) buffers for analyse_p(out=...)."""
        n_mb = (self.width // 16) * (self.height // 16)
        mk = pinned_array if pinned else (lambda s, d: np.zeros(s, dtype=d))
        return mk((n_mb,), MB_OUT_DTYPE), mk((n_mb, self.log_stride), LOG_ENTRY_DTYPE)

    def analyse_p(self, pass_, ref_slots, ref_pocs, cur_poc, out=None, **kw):
        """pcamv_analyse_p: returns (mb records [n_mb], log [n_mb, log_stride])."""
        fi, keep = self._frame_in(pass_, ref_slots, ref_pocs, cur_poc, **kw)
        mbs, log = out if out is not None else self.alloc_outputs()
        self._check(self.lib.pcamv_analyse_p(self.handle, C.byref(fi), _ptr(mbs), _ptr(log)))
        return mbs, log

    def analyse_p_begin(self, pass_, ref_slots, ref_pocs, cur_poc, out, **kw):
        """pcamv_analyse_p_begin: the launch is in flight on return; `out` = page-locked (mbs, log) from alloc_outputs(pinned=True),
        filled row by row by analyse_p_rows."""
        fi, keep = self._frame_in(pass_, ref_slots, ref_pocs, cur_poc, **kw)
        self._check(self.lib.pcamv_analyse_p_begin(self.handle, C.byref(fi), _ptr(out[0]), _ptr(out[1])))

    def analyse_p_rows(self, row):
        """pcamv_analyse_p_rows: blocks until macroblock rows 0..row are in the buffers given to analyse_p_begin; returns the
        number of complete rows."""
        n = C.c_int()
        self._check(self.lib.pcamv_analyse_p_rows(self.handle, int(row), C.byref(n)))
        return int(n.value)

    def frame_upload(self, pass_, ref_slots, ref_pocs, cur_poc, **kw):
        fi, keep = self._frame_in(pass_, ref_slots, ref_pocs, cur_poc, **kw)
        self._check(self.lib.pcamv_frame_upload(self.handle, C.byref(fi)))

    def frame_run(self, pass_=-1, iters=1, per_kernel=False):
        """Mean device ms of one analysis; with per_kernel also (wavefront ms, cost-table ms)."""
        ms = C.c_float()
        mk = (C.c_float * 2)()
        self._check(self.lib.pcamv_frame_run(self.handle, pass_, iters, C.byref(ms), mk if per_kernel else None))
        return (float(ms.value), float(mk[0]), float(mk[1])) if per_kernel else float(ms.value)

    def frame_download(self, want_log=True):
        n_mb = (self.width // 16) * (self.height // 16)
        mbs = np.zeros(n_mb, dtype=MB_OUT_DTYPE)
        log = np.zeros((n_mb, self.log_stride), dtype=LOG_ENTRY_DTYPE) if want_log else None
        self._check(self.lib.pcamv_frame_download(self.handle, _ptr(mbs), _ptr(log) if want_log else None))
        return mbs, log

    def frame_trace(self, enable=True, fetch=False):
        """Switch per-macroblock timestamps on/off; with fetch, return the [n_mb, 2] ns records of the last traced launch."""
        n_mb = (self.width // 16) * (self.height // 16)
        out = np.zeros((n_mb, 2), dtype=np.uint64) if fetch else None
        self._check(self.lib.pcamv_frame_trace(self.handle, int(enable), _ptr(out) if fetch else None))
        return out

    def set_pass2_elide(self, on):
        self._check(self.lib.pcamv_set_pass2_elide(self.handle, int(bool(on))))

    def set_conformant(self, on):
        self._check(self.lib.pcamv_set_conformant(self.handle, int(bool(on))))

    def int_peak_gops(self):
        g = C.c_double()
        self._check(self.lib.pcamv_int_peak(self.handle, C.byref(g)))
        return float(g.value)

    def launch_count(self):
        return int(self.lib.pcamv_launch_count(self.handle))


# -- multi-context launches (pcamv_analyse_p_batch / pcamv_frame_run_batch) -------------------------------------------
def analyse_p_batch(ctxs, frame_args, outs=None):
    """frame_args[i] = (pass_, ref_slots, ref_pocs, cur_poc, kwargs) for ctxs[i]; one wavefront launch for all of them.
    outs: optional preallocated [(mbs, log), ...] (ctx.alloc_outputs).  Returns [(mbs, log), ...]."""
    n = len(ctxs)
    lib = ctxs[0].lib
    fins, keeps = [], []
    if outs is None:
        outs = [c.alloc_outputs() for c in ctxs]
    for c, (pass_, slots, pocs, cur_poc, kw) in zip(ctxs, frame_args):
        fi, keep = c._frame_in(pass_, slots, pocs, cur_poc, **kw)
        fins.append(fi); keeps.append(keep)
    h = (C.c_void_p * n)(*[c.handle for c in ctxs])
    pin = (C.POINTER(FrameIn) * n)(*[C.pointer(f) for f in fins])
    pm = (C.c_void_p * n)(*[o[0].ctypes.data for o in outs])
    pl = (C.c_void_p * n)(*[o[1].ctypes.data for o in outs])
    if lib.pcamv_analyse_p_batch(h, pin, n, pm, pl) != 0:
        raise PcamvError(lib.pcamv_last_error(ctxs[0].handle).decode())
    return outs


def analyse_p_batch_begin(ctxs, frame_args, outs):
    """pcamv_analyse_p_batch_begin: as analyse_p_batch, but returns with the launch in flight; outs = page-locked buffers of every
    context, collected with ctx.analyse_p_rows()."""
    n = len(ctxs)
    lib = ctxs[0].lib
    fins, keeps = [], []
    for c, (pass_, slots, pocs, cur_poc, kw) in zip(ctxs, frame_args):
        fi, keep = c._frame_in(pass_, slots, pocs, cur_poc, **kw)
        fins.append(fi); keeps.append(keep)
    h = (C.c_void_p * n)(*[c.handle for c in ctxs])
    pin = (C.POINTER(FrameIn) * n)(*[C.pointer(f) for f in fins])
    pm = (C.c_void_p * n)(*[o[0].ctypes.data for o in outs])
    pl = (C.c_void_p * n)(*[o[1].ctypes.data for o in outs])
    if lib.pcamv_analyse_p_batch_begin(h, pin, n, pm, pl) != 0:
        raise PcamvError(lib.pcamv_last_error(ctxs[0].handle).decode())


def frame_run_batch(ctxs, pass_, iters=1):
    """Re-launch the staged frames of all contexts in one wavefront (+ cost-table) launch.
    Returns (ms per launch pair, wavefront ms, cost-table ms)."""
    n = len(ctxs)
    lib = ctxs[0].lib
    h = (C.c_void_p * n)(*[c.handle for c in ctxs])
    ms = C.c_float()
    mk = (C.c_float * 2)()
    if lib.pcamv_frame_run_batch(h, n, pass_, iters, C.byref(ms), mk) != 0:
        raise PcamvError(lib.pcamv_last_error(ctxs[0].handle).decode())
    return float(ms.value), float(mk[0]), float(mk[1])
