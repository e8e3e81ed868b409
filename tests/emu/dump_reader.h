// Reader for the tagged binary dumps written by oracle/ref_hooks.c (test infrastructure).
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

struct DumpRec { char tag[5]; uint32_t size; const uint8_t *data; };

struct CallRec      // mirrors pcamv_call_rec_t in oracle/ref_hooks.c
{
    int32_t frame, pass, mb_xy, mb_x, mb_y;
    int32_t i_pixel, i_ref, xoff, yoff, i_ref_cost;
    int32_t me_method, me_range, subme, b_chroma_me, qp;
    int32_t mv_min_fpel[2], mv_max_fpel[2], mv_min_spel[2], mv_max_spel[2];
    int32_t i_mvc, has_thresh, thresh_in, thresh_out;
    int16_t mvp[2];
    int16_t mvc[10][2];
    int16_t mv_in[2]; int32_t cost_in, cost_mv_in;
    int16_t mv[2]; int32_t cost, cost_mv;
    int32_t n_cand, t_ns, pix_sad, pix_satd;
};

struct SliceHdr
{
    int32_t frame, pass, type, qp, nref, with_planes, stride_y, stride_c, lines_y, lines_c, width, fenc_frame, poc, first_mb, last_mb, pad;
};

struct Dump
{
    std::vector<uint8_t> buf;
    std::vector<DumpRec> recs;
    int32_t cfg[24];
    bool load(const char *path)
    {
        FILE *f = fopen(path, "rb");
        if (!f) return false;
        fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
        buf.resize(n);
        if (fread(buf.data(), 1, n, f) != (size_t)n) { fclose(f); return false; }
        fclose(f);
        size_t p = 0;
        memset(cfg, 0, sizeof(cfg));
        while (p + 8 <= buf.size())
        {
            DumpRec r; memcpy(r.tag, &buf[p], 4); r.tag[4] = 0; memcpy(&r.size, &buf[p + 4], 4);
            r.data = &buf[p + 8];
            if (p + 8 + r.size > buf.size()) break;
            recs.push_back(r);
            if (!strcmp(r.tag, "CFG0")) memcpy(cfg, r.data, sizeof(cfg));
            p += 8 + r.size;
        }
        return true;
    }
    const DumpRec *find(const char *tag, size_t from = 0) const
    {
        for (size_t i = from; i < recs.size(); i++)
            if (!strcmp(recs[i].tag, tag)) return &recs[i];
        return nullptr;
    }
};

// planes of one P slice as dumped in a 'SLCB' record
struct SlicePlanes
{
    SliceHdr hd;
    const uint8_t *fenc[3];
    struct Ref { int32_t poc, frame; const uint8_t *y[4]; const uint8_t *u, *v; };   // y[k] -> pixel (0,0)
    std::vector<Ref> refs;
    void parse(const DumpRec &r)
    {
        memcpy(&hd, r.data, sizeof(hd));
        refs.clear();
        if (!hd.with_planes) return;
        const uint8_t *p = r.data + sizeof(hd);
        fenc[0] = p; p += (size_t)hd.stride_y * hd.lines_y;
        fenc[1] = p; p += (size_t)hd.stride_c * hd.lines_c;
        fenc[2] = p; p += (size_t)hd.stride_c * hd.lines_c;
        const size_t luma_sz = (size_t)hd.stride_y * (hd.lines_y + 64), chroma_sz = (size_t)hd.stride_c * (hd.lines_c + 32);
        for (int i = 0; i < hd.nref; i++)
        {
            Ref rf; memcpy(&rf.poc, p, 4); memcpy(&rf.frame, p + 4, 4); p += 8;
            for (int k = 0; k < 4; k++) { rf.y[k] = p + (size_t)hd.stride_y * 32 + 32; p += luma_sz; }
            rf.u = p + (size_t)hd.stride_c * 16 + 16; p += chroma_sz;
            rf.v = p + (size_t)hd.stride_c * 16 + 16; p += chroma_sz;
            refs.push_back(rf);
        }
    }
};
