#ifndef FVENS_B200_PETSC_LITE_TIME
#define FVENS_B200_PETSC_LITE_TIME
#include <petscvec.h>
#include <chrono>
typedef double PetscLogDouble;
static inline PetscErrorCode PetscTime(PetscLogDouble *t) { *t = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); return 0; }
#endif
