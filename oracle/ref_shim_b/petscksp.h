#ifndef FVENS_B200_PETSC_LITE_KSP
#define FVENS_B200_PETSC_LITE_KSP
#include <petscmat.h>
typedef struct _p_KSP* KSP;
typedef struct _p_PC* PC;
#endif
