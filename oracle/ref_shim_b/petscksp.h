#ifndef FVENS_B200_PETSC_LITE_KSP
#define FVENS_B200_PETSC_LITE_KSP
#include <petscmat.h>
typedef struct _p_KSP* KSP;
typedef struct _p_PC* PC;
static inline PetscErrorCode KSPSolve(KSP, Vec, Vec) { return PETSC_ERR_SUP; }
static inline PetscErrorCode KSPGetIterationNumber(KSP, PetscInt *n) { *n = 0; return PETSC_ERR_SUP; }
static inline PetscErrorCode KSPGetPC(KSP, PC *pc) { *pc = NULL; return PETSC_ERR_SUP; }
static inline PetscErrorCode KSPGetOperators(KSP, Mat *a, Mat *b) { if(a) *a = NULL; if(b) *b = NULL; return PETSC_ERR_SUP; }
static inline PetscErrorCode PCGAMGSetReuseInterpolation(PC, PetscBool) { return PETSC_ERR_SUP; }
#endif
