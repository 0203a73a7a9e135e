/*
 * topology.c -- ORACLE (test infrastructure, see oracle.h): per-atom masses
 * from a .pdb / .gro topology.
 *
 * Restates what read_tps_conf(..., bMass=TRUE) gives the reference at
 * knn_rms.cpp:150-153,181-182: PDB/GRO files carry no masses, GROMACS looks
 * them up by atom NAME in share/top/atommass.dat (longest case-sensitive
 * prefix; so "CA" in a protein is carbon, not "Ca" calcium).  Values are
 * the GROMACS 4.5+/5.x table; unknown names get 12.011.
 */
#include "oracle.h"
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static const struct { const char *name; float mass; } mass_table[] = {
    /* two-letter, case-sensitive entries are tried first */
    {"Cl", 35.45300f}, {"Br", 79.90000f}, {"Na", 22.98970f}, {"Mg", 24.30500f}, {"Ca", 40.08000f},
    {"Fe", 55.84700f}, {"Zn", 65.37000f}, {"Cu", 63.54600f}, {"Si", 28.08000f}, {"Al", 26.98150f},
    {"H", 1.00790f},   {"C", 12.01070f},  {"N", 14.00670f},  {"O", 15.99940f},  {"S", 32.06500f},
    {"P", 30.97380f},  {"F", 18.99840f},  {"B", 10.81100f},  {"I", 126.90450f}, {"K", 39.10200f},
};

static float mass_for_name(const char *raw)
{
    char name[16];
    int n = 0;
    /* drop blanks and a leading digit run ("1HB" -> "HB") */
    while (*raw && (isspace((unsigned char)*raw) || isdigit((unsigned char)*raw))) raw++;
    while (*raw && !isspace((unsigned char)*raw) && n < 15) name[n++] = *raw++;
    name[n] = 0;
    size_t best = 0;
    float m = 12.011f;
    for (size_t i = 0; i < sizeof(mass_table) / sizeof(mass_table[0]); i++) {
        size_t l = strlen(mass_table[i].name);
        if (l > best && strncmp(name, mass_table[i].name, l) == 0) { best = l; m = mass_table[i].mass; }
    }
    return m;
}

static int has_suffix(const char *s, const char *suf)
{
    size_t a = strlen(s), b = strlen(suf);
    return a >= b && strcmp(s + a - b, suf) == 0;
}

int oracle_top_masses(const char *path, float *mass, int max_atoms)
{
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    char line[512];
    int n = 0;
    if (has_suffix(path, ".gro")) {
        int declared = 0;
        if (!fgets(line, sizeof line, f) || !fgets(line, sizeof line, f)) { fclose(f); return -2; }
        declared = atoi(line);
        for (int i = 0; i < declared && fgets(line, sizeof line, f); i++) {
            if (strlen(line) < 15) break;
            char nm[6];
            memcpy(nm, line + 10, 5);
            nm[5] = 0;
            if (n < max_atoms) mass[n] = mass_for_name(nm);
            n++;
        }
    } else {
        while (fgets(line, sizeof line, f)) {
            if (strncmp(line, "ENDMDL", 6) == 0) break; /* first model only */
            if (strncmp(line, "ATOM", 4) != 0 && strncmp(line, "HETATM", 6) != 0) continue;
            if (strlen(line) < 16) continue;
            char nm[5];
            memcpy(nm, line + 12, 4);
            nm[4] = 0;
            if (n < max_atoms) mass[n] = mass_for_name(nm);
            n++;
        }
    }
    fclose(f);
    return n;
}
