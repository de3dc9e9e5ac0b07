// Numeric tables embedded into the shared object (no file IO at run time).
//   rb_sobol_table : Joe-Kuo direction matrices, 1024 dims x 52 x uint64 (reference table src/sobol.inc:32-35)
//   rb_ltc_table   : LTC matrices fitted to Blinn-Phong, 128 x 128 x 9 float (reference table src/ltc.inc:14)
// Both are produced by tools/extract_tables.py; RB_DATA_DIR is set by the build (redner_b200/build.py).
#ifndef RB_DATA_DIR
#error "RB_DATA_DIR must be defined"
#endif
#define RB_STR2(x) #x
#define RB_STR(x) RB_STR2(x)
__asm__(".section .rodata\n"
        ".balign 16\n"
        ".global rb_sobol_table_begin\n"
        "rb_sobol_table_begin:\n"
        ".incbin \"" RB_STR(RB_DATA_DIR) "/sobol_joe_kuo_1024x52_u64.bin\"\n"
        ".global rb_sobol_table_end\n"
        "rb_sobol_table_end:\n"
        ".balign 16\n"
        ".global rb_ltc_table_begin\n"
        "rb_ltc_table_begin:\n"
        ".incbin \"" RB_STR(RB_DATA_DIR) "/ltc_blinn_phong_128x128x9_f32.bin\"\n"
        ".global rb_ltc_table_end\n"
        "rb_ltc_table_end:\n"
        ".section .text\n");
