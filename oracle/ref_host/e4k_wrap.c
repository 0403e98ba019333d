/*
 * e4k_wrap.c -- compiles the UNMODIFIED reference RTL/Src/tuner_e4k.c for the host and, from inside
 * the same translation unit, exposes the file-static selection routines the front-end known-answer
 * tests need (they cannot be reached from another object file).
 *
 * TEST INFRASTRUCTURE ONLY (oracle A).  No reference source is copied: the file is #included from
 * where it lies (REF_TUNER_E4K_C is passed by oracle/Makefile as an absolute path).
 *
 *   ref_e4k_rf_filter   -> choose_rf_filter   tuner_e4k.c:250-277 (closest_arr_idx :231-247)
 *   ref_e4k_if_bw_index -> find_if_bw         tuner_e4k.c:363-372
 *   ref_e4k_if_bw_hz    -> if_filter_bw[][]   tuner_e4k.c:165-203 (the three bandwidth tables)
 */
#include REF_TUNER_E4K_C

int ref_e4k_rf_filter(int band, uint32_t freq) { return choose_rf_filter((enum e4k_band)band, freq); }

int ref_e4k_if_bw_index(int filter, uint32_t bw) { return find_if_bw((enum e4k_if_filter)filter, bw); }

/* 0 when (filter, idx) is outside the tables */
uint32_t ref_e4k_if_bw_hz(int filter, int idx)
{
    if (filter < 0 || filter >= (int)ARRAY_SIZE(if_filter_bw)) return 0;
    if (idx < 0 || idx >= (int)if_filter_bw_len[filter]) return 0;
    return if_filter_bw[filter][idx];
}
