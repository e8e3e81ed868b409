import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import pcamv_loader, refrun
pcamv = pcamv_loader.load()
name = sys.argv[1] if len(sys.argv) > 1 else "qcif_hex5"
dump = pcamv.dumpfmt.Dump(refrun.golden_dump_path(name))
calls, refine = dump.calls()
s = next(x for x in dump.slices() if x.with_planes)
c = dump.cfg
ctx = pcamv.PcamvContext(s.width, s.lines_y, me_method=c["me_method"], me_range=c["me_range"], subpel_refine=c["subme"],
                         chroma_me=c["chroma_me"], max_refs=c["refs"], mv_range=c["mv_range"])
t = dump.cost_tables[s.qp]
ctx.set_qp_tables(s.qp, t["lambda"], t["cost_mv"], t["cost_ref"])
ctx.put_fenc(s.fenc[0][:, :s.width], s.fenc[1][:, :s.width // 2], s.fenc[2][:, :s.width // 2])
for slot, r in enumerate(s.refs):
    ctx.put_ref_planes(slot, r["poc"], r["luma"], r["u"], r["v"])
sel = calls["frame"] == s.frame
rc, rf = calls[sel], refine[sel]
res = ctx.me_search_batch(pcamv.dumpfmt.calls_to_abi(rc, rf))
ok = (res["mv"] == rc["mv"]).all(axis=1) & (res["cost"] == rc["cost"])
print("total", len(rc), "bad", int((~ok).sum()))
for mode in (0, 1):
    for pix in range(7):
        m = (rf == bool(mode)) & (rc["i_pixel"] == pix)
        if m.sum():
            print("mode", mode, "pix", pix, "n", int(m.sum()), "bad", int((~ok & m).sum()))
idx = np.nonzero(~ok)[0][:12]
for i in idx:
    print(i, "mode", int(rf[i]), "pix", rc["i_pixel"][i], "mb", rc["mb_x"][i], rc["mb_y"][i], "off", rc["xoff"][i], rc["yoff"][i],
          "mvp", rc["mvp"][i], "got", res["mv"][i], res["cost"][i], res["cost_mv"][i], "want", rc["mv"][i], rc["cost"][i], rc["cost_mv"][i])
