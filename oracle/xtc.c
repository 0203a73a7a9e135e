/*
 * xtc.c -- ORACLE (test infrastructure, see oracle.h): sequential XTC decoder.
 *
 * Stands in for GROMACS open_xtc / read_first_xtc / read_next_xtc as the
 * reference calls them (knn_rms.cpp:155-157,186-203; wrappers mdsctk.cpp:217-267).
 * libgromacs is not vendored in the reference, so this restates the published
 * xtc "xdr3dfcoord" compressed-coordinate format.
 * Written as a plain bit reader over an in-memory copy of the file.
 */
#include "oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#define XTC_MAGIC 1995
#define FIRSTIDX 9

static const int magicints[] = {
    0, 0, 0, 0, 0, 0, 0, 0, 0, 8, 10, 12, 16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406,
    512, 645, 812, 1024, 1290, 1625, 2048, 2580, 3250, 4096, 5060, 6501, 8192, 10321, 13003, 16384, 20642,
    26007, 32768, 41285, 52015, 65536, 82570, 104031, 131072, 165140, 208063, 262144, 330280, 416127,
    524287, 660561, 832255, 1048576, 1321122, 1664510, 2097152, 2642245, 3329021, 4194304, 5284491,
    6658042, 8388607, 10568983, 13316085, 16777216};
#define LASTIDX ((int)(sizeof(magicints) / sizeof(magicints[0])))

typedef struct {
    const unsigned char *p;
    size_t nbytes;
    size_t bitpos;
} bitreader;

static unsigned read_bits(bitreader *br, int nbits)
{
    unsigned v = 0;
    for (int i = 0; i < nbits; i++) {
        size_t byte = br->bitpos >> 3;
        unsigned bit = 0;
        if (byte < br->nbytes) bit = (br->p[byte] >> (7 - (br->bitpos & 7))) & 1u;
        v = (v << 1) | bit;
        br->bitpos++;
    }
    return v;
}

/* Read `nbits` as a little-endian-by-byte big number and split it into three
 * mixed-radix digits (sizes[0] most significant). */
static void read_triple(bitreader *br, int nbits, const unsigned sizes[3], int out[3])
{
    unsigned __int128 v = 0;
    int shift = 0;
    while (nbits > 8) {
        v |= (unsigned __int128)read_bits(br, 8) << shift;
        shift += 8;
        nbits -= 8;
    }
    if (nbits > 0) v |= (unsigned __int128)read_bits(br, nbits) << shift;
    out[2] = (int)(v % sizes[2]);
    v /= sizes[2];
    out[1] = (int)(v % sizes[1]);
    out[0] = (int)(v / sizes[1]);
}

static int bit_length(unsigned long long x)
{
    int n = 0;
    while (x) { n++; x >>= 1; }
    return n;
}

static int bit_length_product3(const unsigned s[3])
{
    unsigned __int128 p = (unsigned __int128)s[0] * s[1] * s[2];
    int n = 0;
    while (p) { n++; p >>= 1; }
    return n;
}

static int32_t be_i32(const unsigned char *p)
{
    return (int32_t)(((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3]);
}
static float be_f32(const unsigned char *p)
{
    uint32_t u = (uint32_t)be_i32(p);
    float f;
    memcpy(&f, &u, 4);
    return f;
}

static unsigned char *slurp(const char *path, size_t *len)
{
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    unsigned char *buf = (unsigned char *)malloc(n > 0 ? (size_t)n : 1);
    if (buf && n > 0 && fread(buf, 1, (size_t)n, f) != (size_t)n) { free(buf); buf = NULL; }
    fclose(f);
    *len = (size_t)(n > 0 ? n : 0);
    return buf;
}

/* Decode one frame starting at buf+*off; advances *off.  out may be NULL
 * (header walk only).  Returns natoms or <0 on error. */
static int decode_frame(const unsigned char *buf, size_t len, size_t *off, float *out, int expect_atoms)
{
    size_t o = *off;
    if (o + 56 + 4 > len) return -1;
    if (be_i32(buf + o) != XTC_MAGIC) return -2;
    int natoms = be_i32(buf + o + 4);
    /* step at +8, time at +12, box 9 floats at +16 */
    o += 16 + 36;
    int lsize = be_i32(buf + o);
    o += 4;
    if (lsize != natoms || natoms <= 0) return -3;
    if (expect_atoms > 0 && natoms != expect_atoms) return -4;

    if (natoms <= 9) {
        if (o + (size_t)natoms * 12 > len) return -1;
        if (out)
            for (int i = 0; i < natoms * 3; i++) out[i] = be_f32(buf + o + 4 * (size_t)i);
        o += (size_t)natoms * 12;
        *off = o;
        return natoms;
    }

    if (o + 4 + 24 + 4 + 4 > len) return -1;
    float precision = be_f32(buf + o);
    o += 4;
    int minint[3], maxint[3];
    for (int d = 0; d < 3; d++) minint[d] = be_i32(buf + o + 4 * d);
    o += 12;
    for (int d = 0; d < 3; d++) maxint[d] = be_i32(buf + o + 4 * d);
    o += 12;
    int smallidx = be_i32(buf + o);
    o += 4;
    int nbytes = be_i32(buf + o);
    o += 4;
    if (nbytes < 0 || o + (size_t)nbytes > len) return -1;
    size_t padded = ((size_t)nbytes + 3) & ~(size_t)3;

    if (out) {
        unsigned sizeint[3], sizesmall[3];
        int bitsizeint[3] = {0, 0, 0};
        int bitsize;
        for (int d = 0; d < 3; d++) sizeint[d] = (unsigned)(maxint[d] - minint[d] + 1);
        if ((sizeint[0] | sizeint[1] | sizeint[2]) > 0xffffffu) {
            for (int d = 0; d < 3; d++) bitsizeint[d] = bit_length(sizeint[d]);
            bitsize = 0;
        } else {
            bitsize = bit_length_product3(sizeint);
        }
        if (smallidx < FIRSTIDX || smallidx >= LASTIDX) return -5;
        int tmpidx = smallidx - 1;
        if (tmpidx < FIRSTIDX) tmpidx = FIRSTIDX;
        int smaller = magicints[tmpidx] / 2;
        int smallnum = magicints[smallidx] / 2;
        sizesmall[0] = sizesmall[1] = sizesmall[2] = (unsigned)magicints[smallidx];

        const float inv = 1.0f / precision;
        bitreader br = {buf + o, (size_t)nbytes, 0};
        int run = 0, i = 0, written = 0;
        int thisc[3], prevc[3];
        while (i < natoms) {
            if (bitsize == 0) {
                for (int d = 0; d < 3; d++) thisc[d] = (int)read_bits(&br, bitsizeint[d]);
            } else {
                read_triple(&br, bitsize, sizeint, thisc);
            }
            i++;
            for (int d = 0; d < 3; d++) { thisc[d] += minint[d]; prevc[d] = thisc[d]; }

            int flag = (int)read_bits(&br, 1);
            int is_smaller = 0;
            if (flag) {
                run = (int)read_bits(&br, 5);
                is_smaller = run % 3;
                run -= is_smaller;
                is_smaller--;
            }
            if (run > 0) {
                for (int k = 0; k < run; k += 3) {
                    if (written >= natoms) return -6;
                    read_triple(&br, smallidx, sizesmall, thisc);
                    i++;
                    for (int d = 0; d < 3; d++) thisc[d] += prevc[d] - smallnum;
                    if (k == 0) {
                        /* the second atom of the group is stored ahead of the first */
                        for (int d = 0; d < 3; d++) { int t = thisc[d]; thisc[d] = prevc[d]; prevc[d] = t; }
                        for (int d = 0; d < 3; d++) out[3 * written + d] = (float)prevc[d] * inv;
                        written++;
                    } else {
                        for (int d = 0; d < 3; d++) prevc[d] = thisc[d];
                    }
                    if (written >= natoms) return -6;
                    for (int d = 0; d < 3; d++) out[3 * written + d] = (float)thisc[d] * inv;
                    written++;
                }
            } else {
                if (written >= natoms) return -6;
                for (int d = 0; d < 3; d++) out[3 * written + d] = (float)thisc[d] * inv;
                written++;
            }
            smallidx += is_smaller;
            if (smallidx < FIRSTIDX || smallidx >= LASTIDX) return -5;
            if (is_smaller < 0) {
                smallnum = smaller;
                smaller = (smallidx > FIRSTIDX) ? magicints[smallidx - 1] / 2 : 0;
            } else if (is_smaller > 0) {
                smaller = smallnum;
                smallnum = magicints[smallidx] / 2;
            }
            sizesmall[0] = sizesmall[1] = sizesmall[2] = (unsigned)magicints[smallidx];
        }
        if (written != natoms) return -6;
    }
    o += padded;
    if (o > len) o = len; /* last frame may lack padding */
    *off = o;
    return natoms;
}

int oracle_xtc_scan(const char *path, int *natoms, long long *nframes)
{
    size_t len;
    unsigned char *buf = slurp(path, &len);
    if (!buf) return -1;
    size_t off = 0;
    long long n = 0;
    int na = 0;
    while (off + 60 <= len) {
        int r = decode_frame(buf, len, &off, NULL, na);
        if (r < 0) break;
        na = r;
        n++;
    }
    free(buf);
    *natoms = na;
    *nframes = n;
    return n > 0 ? 0 : -2;
}

int oracle_xtc_read(const char *path, int natoms, long long nframes, float *xyz)
{
    size_t len;
    unsigned char *buf = slurp(path, &len);
    if (!buf) return -1;
    size_t off = 0;
    for (long long f = 0; f < nframes; f++) {
        int r = decode_frame(buf, len, &off, xyz + (size_t)f * natoms * 3, natoms);
        if (r < 0) { free(buf); return r; }
    }
    free(buf);
    return 0;
}
