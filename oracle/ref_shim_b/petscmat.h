/* Stand-in for <petscmat.h>: the residual path only names the Mat type (Jacobian assembly is declared in the same
 * classes but never called by the harness). TEST INFRASTRUCTURE ONLY. */
#ifndef FVENS_B200_PETSC_LITE_MAT
#define FVENS_B200_PETSC_LITE_MAT
#include <petscvec.h>
typedef struct _p_Mat* Mat;
static inline PetscErrorCode MatSetValuesBlocked(Mat, PetscInt, const PetscInt*, PetscInt, const PetscInt*, const PetscScalar*, InsertMode) { return PETSC_ERR_SUP; }
#endif
