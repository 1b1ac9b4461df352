/* Stand-in for <petscmat.h>: the residual path only names the Mat type (Jacobian assembly is declared in the same
 * classes but never called by the harness). TEST INFRASTRUCTURE ONLY. */
#ifndef FVENS_B200_PETSC_LITE_MAT
#define FVENS_B200_PETSC_LITE_MAT
#include <petscvec.h>
typedef struct _p_Mat* Mat;
typedef enum { MAT_FLUSH_ASSEMBLY = 1, MAT_FINAL_ASSEMBLY = 0 } MatAssemblyType;
typedef enum { MAT_NEW_NONZERO_ALLOCATION_ERR = 19, MAT_USE_HASH_TABLE = 5, MAT_NEW_NONZERO_LOCATIONS = 2 } MatOption;
/* the implicit solvers' calls: never reached by the harness (explicit pseudo-time only) */
static inline PetscErrorCode MatZeroEntries(Mat) { return PETSC_ERR_SUP; }
static inline PetscErrorCode MatView(Mat, PetscViewer) { return PETSC_ERR_SUP; }
static inline PetscErrorCode MatShellGetContext(Mat, void*) { return PETSC_ERR_SUP; }
static inline PetscErrorCode MatSetValues(Mat, PetscInt, const PetscInt*, PetscInt, const PetscInt*, const PetscScalar*, InsertMode) { return PETSC_ERR_SUP; }
static inline PetscErrorCode MatSetOption(Mat, MatOption, PetscBool) { return PETSC_ERR_SUP; }
static inline PetscErrorCode MatAssemblyBegin(Mat, MatAssemblyType) { return PETSC_ERR_SUP; }
static inline PetscErrorCode MatAssemblyEnd(Mat, MatAssemblyType) { return PETSC_ERR_SUP; }
static inline PetscErrorCode MatSetValuesBlocked(Mat, PetscInt, const PetscInt*, PetscInt, const PetscInt*, const PetscScalar*, InsertMode) { return PETSC_ERR_SUP; }
#endif
