"""Parser for the tagged binary dumps written by the instrumented reference (oracle/ref_hooks.c).

Record stream: 4-byte tag, uint32 size, payload.  Used by tests/ and bench.py to turn recorded
x264_me_search_ref / x264_me_refine_qpel calls and frame planes into C-ABI inputs and expected outputs.
"""
import numpy as np

from .host import ME_CALL_DTYPE

CALL_REC_DTYPE = np.dtype([
    ("frame", "<i4"), ("pass_", "<i4"), ("mb_xy", "<i4"), ("mb_x", "<i4"), ("mb_y", "<i4"),
    ("i_pixel", "<i4"), ("i_ref", "<i4"), ("xoff", "<i4"), ("yoff", "<i4"), ("i_ref_cost", "<i4"),
    ("me_method", "<i4"), ("me_range", "<i4"), ("subme", "<i4"), ("b_chroma_me", "<i4"), ("qp", "<i4"),
    ("mv_min_fpel", "<i4", 2), ("mv_max_fpel", "<i4", 2), ("mv_min_spel", "<i4", 2), ("mv_max_spel", "<i4", 2),
    ("i_mvc", "<i4"), ("has_thresh", "<i4"), ("thresh_in", "<i4"), ("thresh_out", "<i4"),
    ("mvp", "<i2", 2), ("mvc", "<i2", (10, 2)), ("mv_in", "<i2", 2), ("cost_in", "<i4"), ("cost_mv_in", "<i4"),
    ("mv", "<i2", 2), ("cost", "<i4"), ("cost_mv", "<i4"), ("n_cand", "<i4"), ("t_ns", "<i4"),
    ("pix_sad", "<i4"), ("pix_satd", "<i4")], align=True)
assert CALL_REC_DTYPE.itemsize == 192

MBAN_DTYPE = np.dtype([
    ("frame", "<i4"), ("pass_", "<i4"), ("mb_xy", "<i4"), ("type", "<i4"), ("partition", "<i4"),
    ("sub", "<i4", 4), ("b_skip_mc", "<i4"), ("qp", "<i4"),
    ("mv", "<i2", (16, 2)), ("ref", "i1", 4), ("pskip_mv", "<i2", 2)])
assert MBAN_DTYPE.itemsize == 44 + 64 + 4 + 4

EMBD_MB_DTYPE = np.dtype([
    ("type", "<i4"), ("qp", "<i4"), ("partition", "<i4"), ("used", "u1"), ("sub", "u1", 4), ("pad", "u1", 3),
    ("mv_stego", "<i2", (16, 2)), ("inter_stego_cost", "<i4", 16), ("ref", "i1", 16), ("mv", "<i2", (16, 2)),
    ("pskip_mv", "<i2", 2)])
assert EMBD_MB_DTYPE.itemsize == 232

CFG_NAMES = ["width", "height", "mb_w", "mb_h", "me_method", "me_range", "subme", "refs", "chroma_me", "mv_range",
             "qp", "b_cabac", "inter", "mixed_refs", "fast_pskip", "dct_decimate", "keyint", "transform_8x8",
             "trellis", "threads", "b4_stride", "b8_stride", "mb_stride", "emrate_x1e6"]


class Slice:
    """One 'SLCB' record: header + (optionally) fenc and reference planes as numpy views."""

    def __init__(self, payload):
        hd = np.frombuffer(payload, dtype="<i4", count=16)
        (self.frame, self.pass_, self.type, self.qp, self.nref, self.with_planes, self.stride_y, self.stride_c,
         self.lines_y, self.lines_c, self.width, self.fenc_frame, self.poc, self.first_mb, self.last_mb, _) = [int(x) for x in hd]
        self.refs = []
        self.fenc = None
        if not self.with_planes:
            return
        buf = np.frombuffer(payload, dtype=np.uint8)
        p = 64
        sy, sc, ly, lc = self.stride_y, self.stride_c, self.lines_y, self.lines_c

        def take(n, shape):
            nonlocal p
            a = buf[p:p + n].reshape(shape)
            p += n
            return a

        self.fenc = (take(sy * ly, (ly, sy)), take(sc * lc, (lc, sc)), take(sc * lc, (lc, sc)))
        for _ in range(self.nref):
            poc, frame = [int(x) for x in np.frombuffer(payload, dtype="<i4", count=2, offset=p)]
            p += 8
            luma = [take(sy * (ly + 64), (ly + 64, sy)) for _ in range(4)]
            u = take(sc * (lc + 32), (lc + 32, sc))
            v = take(sc * (lc + 32), (lc + 32, sc))
            self.refs.append({"poc": poc, "frame": frame, "luma": luma, "u": u, "v": v})


class Dump:
    def __init__(self, path_or_bytes):
        if isinstance(path_or_bytes, (bytes, bytearray, memoryview)):
            self.raw = bytes(path_or_bytes)
        else:
            with open(path_or_bytes, "rb") as f:
                self.raw = f.read()
        self.records = []           # (tag, offset, size)
        p, n = 0, len(self.raw)
        while p + 8 <= n:
            tag = self.raw[p:p + 4].decode("latin-1")
            size = int.from_bytes(self.raw[p + 4:p + 8], "little")
            if p + 8 + size > n:
                break
            self.records.append((tag, p + 8, size))
            p += 8 + size
        self.cfg = {}
        self.cost_tables = {}       # qp -> dict(lambda, cost_mv[32769], cost_ref[3,33])
        for tag, off, size in self.records:
            if tag == "CFG0":
                vals = np.frombuffer(self.raw, dtype="<i4", count=24, offset=off)
                self.cfg = {k: int(v) for k, v in zip(CFG_NAMES, vals)}
            elif tag == "CMV0":
                qp, lam = [int(x) for x in np.frombuffer(self.raw, dtype="<i4", count=2, offset=off)]
                cm = np.frombuffer(self.raw, dtype="<i2", count=32769, offset=off + 8)
                cr = np.frombuffer(self.raw, dtype="<u2", count=99, offset=off + 8 + 32769 * 2).reshape(3, 33)
                self.cost_tables[qp] = {"lambda": lam, "cost_mv": cm, "cost_ref": cr}

    def payload(self, rec):
        _, off, size = rec
        return memoryview(self.raw)[off:off + size]

    def slices(self):
        for rec in self.records:
            if rec[0] == "SLCB":
                yield Slice(self.payload(rec))

    def _stack(self, tag, dtype):
        recs = [r for r in self.records if r[0] == tag]
        out = np.zeros(len(recs), dtype=dtype)
        for i, r in enumerate(recs):
            out[i] = np.frombuffer(self.raw, dtype=dtype, count=1, offset=r[1])[0]
        return out

    def calls(self):
        """All MESR / MERQ records in file order -> (structured array, is_refine bool array)."""
        recs = [r for r in self.records if r[0] in ("MESR", "MERQ")]
        out = np.zeros(len(recs), dtype=CALL_REC_DTYPE)
        refine = np.zeros(len(recs), dtype=bool)
        for i, r in enumerate(recs):
            out[i] = np.frombuffer(self.raw, dtype=CALL_REC_DTYPE, count=1, offset=r[1])[0]
            refine[i] = r[0] == "MERQ"
        return out, refine

    def slice_units(self):
        """Per P-slice pass with planes, in file order: dict(slice, ctx (SLCX fields), calls, refine, mban, embd).
        `embd` is the EMBD record written at the end of that pass (pass 1 only) or None."""
        units, cur = [], None
        embeds = iter(self.embeds())
        for rec in self.records:
            tag, off, size = rec
            if tag == "SLCB":
                s = Slice(self.payload(rec))
                cur = None
                if s.type == 0:
                    cur = {"slice": s, "ctx": None, "calls": [], "refine": [], "mban": [], "embd": None}
                    units.append(cur)
            elif cur is None:
                if tag == "EMBD":
                    next(embeds)
                continue
            elif tag == "SLCX":
                hd = np.frombuffer(self.raw, dtype="<i4", count=36, offset=off)
                n_mb = int(hd[35])
                cur["ctx"] = {"cur_poc": int(hd[0]), "n_ref": int(hd[1]), "ref_poc": [int(x) for x in hd[2:18]],
                              "col_n_ref": int(hd[18]), "col_inv_ref_poc": [int(x) for x in hd[19:35]], "n_mb": n_mb,
                              "col_ref8": np.frombuffer(self.raw, dtype=np.int8, count=4 * n_mb, offset=off + 144),
                              "col_mv4": np.frombuffer(self.raw, dtype="<i2", count=32 * n_mb, offset=off + 144 + 4 * n_mb)}
            elif tag in ("MESR", "MERQ"):
                cur["calls"].append(np.frombuffer(self.raw, dtype=CALL_REC_DTYPE, count=1, offset=off)[0])
                cur["refine"].append(tag == "MERQ")
            elif tag == "MBAN":
                cur["mban"].append(np.frombuffer(self.raw, dtype=MBAN_DTYPE, count=1, offset=off)[0])
            elif tag == "EMBD":
                cur["embd"] = next(embeds)
            elif tag == "SLCE":
                cur = None
        for u in units:
            u["calls"] = np.array(u["calls"], dtype=CALL_REC_DTYPE)
            u["refine"] = np.array(u["refine"], dtype=bool)
            u["mban"] = np.array(u["mban"], dtype=MBAN_DTYPE)
        return units

    def counters(self):
        """'CNT0' records in file order (one per dumped P-slice pass): the reference's work counters of that pass."""
        names = ["sad", "satd", "ih_luma", "ih_chroma", "pix_sad", "pix_satd", "searches", "refines", "ih_calls"]
        out = []
        for tag, off, size in self.records:
            if tag != "CNT0":
                continue
            c = np.frombuffer(self.raw, dtype="<u8", count=9, offset=off)
            t = np.frombuffer(self.raw, dtype="<f8", count=3, offset=off + 72)
            d = {k: int(v) for k, v in zip(names, c)}
            d.update(t_me=float(t[0]), t_ih=float(t[1]), t_analyse_p=float(t[2]), pix_avg=0, pix_chroma_mc=0, pix_dct=0, ih_pix_avg=0, ih_pix_chroma_mc=0, ih_pix_dct=0, ih_pix_satd=0)
            out.append(d)
        # 'CNT1' (newer dumps) follows its 'CNT0': pixels through quarter-pel averaging, chroma MC, DCT/quant/IDCT
        k = 0
        for tag, off, size in self.records:
            if tag == "CNT0":
                k += 1
            elif tag == "CNT1" and k:
                x = np.frombuffer(self.raw, dtype="<u8", count=7, offset=off)
                out[k - 1].update(pix_avg=int(x[0]), pix_chroma_mc=int(x[1]), pix_dct=int(x[2]), ih_pix_avg=int(x[3]),
                                  ih_pix_chroma_mc=int(x[4]), ih_pix_dct=int(x[5]), ih_pix_satd=int(x[6]))
        return out

    def quant_tables(self):
        """qp -> dict of the 'QNT0' record (quantiser tables of the P slices)."""
        out = {}
        for tag, off, size in self.records:
            if tag != "QNT0":
                continue
            qp, qpc, lam2 = [int(x) for x in np.frombuffer(self.raw, dtype="<i4", count=3, offset=off)]
            u16 = np.frombuffer(self.raw, dtype="<u2", count=64, offset=off + 12)
            dq = np.frombuffer(self.raw, dtype="<i4", count=192, offset=off + 12 + 128)
            out[qp] = {"chroma_qp": qpc, "lambda2_chroma": lam2, "quant4_mf": (u16[0:16], u16[32:48]),
                       "quant4_bias": (u16[16:32], u16[48:64]), "dequant4_mf": (dq[:96], dq[96:])}
        return out

    def mb_decisions(self):
        return self._stack("MBAN", MBAN_DTYPE)

    def embeds(self):
        """List of dicts, one per 'EMBD' record (end of pass 1 of a P frame)."""
        out = []
        for tag, off, size in self.records:
            if tag != "EMBD":
                continue
            frame, n_mb, length, an, num_filp = [int(x) for x in np.frombuffer(self.raw, dtype="<i4", count=5, offset=off)]
            p = off + 20
            mbs = np.frombuffer(self.raw, dtype=EMBD_MB_DTYPE, count=n_mb, offset=p); p += 232 * n_mb
            cover = np.frombuffer(self.raw, dtype=np.uint8, count=length, offset=p); p += length
            rho = np.frombuffer(self.raw, dtype="<f4", count=length, offset=p); p += 4 * length
            msg = np.frombuffer(self.raw, dtype=np.uint8, count=max(an, 0), offset=p); p += max(an, 0)
            stego = np.frombuffer(self.raw, dtype=np.uint8, count=length, offset=p); p += length
            filp = np.frombuffer(self.raw, dtype=np.int8, count=length, offset=p); p += length
            out.append({"frame": frame, "n_mb": n_mb, "length": length, "an": an, "num_filp": num_filp, "mbs": mbs,
                        "cover": cover, "rho": rho, "message": msg, "stego": stego, "filp": filp})
        return out


def calls_to_abi(recs, refine, slot_of_ref=None):
    """Convert recorded calls into pcamv_me_call records (include/pcamv.h)."""
    out = np.zeros(len(recs), dtype=ME_CALL_DTYPE)
    out["mode"] = refine.astype(np.int32)
    for k in ("mb_x", "mb_y", "xoff", "yoff", "i_pixel", "i_ref_cost", "mv_min_fpel", "mv_max_fpel", "mv_min_spel",
              "mv_max_spel", "has_thresh", "thresh_in", "mvp", "mvc", "mv_in", "cost_in", "cost_mv_in"):
        out[k] = recs[k]
    out["i_mvc"] = np.minimum(recs["i_mvc"], 10)
    out["ref_slot"] = recs["i_ref"] if slot_of_ref is None else np.asarray(slot_of_ref)[recs["i_ref"]]
    return out
