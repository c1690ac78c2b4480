/* oracle/ref_blob.S -- TEST INFRASTRUCTURE ONLY.
 * Embeds the reference's precompiled kernel blob and LZ4 dictionary from where
 * they lie (paths passed by oracle/Makefile, normally under /root/reference) under the symbol names that the
 * reference's generated kernels_75.c / kernels_dict.c would define
 * (CMakeLists.txt:93-111, resources/kernels.h:16-24). */
    .section .rodata
    .global kernels_75
    .type kernels_75, @object
    .balign 16
kernels_75:
    .incbin REF_KERNELS_75
    .size kernels_75, . - kernels_75
    .global kernels_dict
    .type kernels_dict, @object
    .balign 16
kernels_dict:
    .incbin REF_KERNELS_DICT
    .size kernels_dict, . - kernels_dict
    .section .note.GNU-stack,"",@progbits
