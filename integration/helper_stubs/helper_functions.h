/* Empty stand-in for the CUDA-samples helper_functions.h (nothing from it is used by lyap_calculate.cu). */
