/* Embeds data/lab_lut_s16.bin (see tools/make_lab_lut.py) into libmosaic_b200.so. */
    .section .rodata
    .global mm_lab_lut_s16
    .global mm_lab_lut_s16_end
    .balign 16
mm_lab_lut_s16:
    .incbin "../data/lab_lut_s16.bin"
mm_lab_lut_s16_end:
    .section .note.GNU-stack,"",@progbits
