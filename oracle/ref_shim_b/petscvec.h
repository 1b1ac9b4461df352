/* Serial stand-in for the PETSc Vec interface used by the reference's spatial sources (ghosted block Vecs as
 * storage; no arithmetic lives in PETSc on this path). TEST INFRASTRUCTURE ONLY (oracle/ref_tier_c.cpp). */
#ifndef FVENS_B200_PETSC_LITE_VEC
#define FVENS_B200_PETSC_LITE_VEC
#include <mpi.h>
#include <cstddef>
#include <vector>

typedef int PetscErrorCode;
typedef int PetscInt;
typedef double PetscScalar;
typedef double PetscReal;
typedef int PetscMPIInt;
typedef enum { PETSC_FALSE = 0, PETSC_TRUE = 1 } PetscBool;
typedef enum { NOT_SET_VALUES, INSERT_VALUES, ADD_VALUES } InsertMode;
typedef enum { SCATTER_FORWARD = 0, SCATTER_REVERSE = 1 } ScatterMode;
typedef enum { NORM_1 = 0, NORM_2 = 1, NORM_INFINITY = 3 } NormType;
#define PETSC_COMM_WORLD MPI_COMM_WORLD
#define PETSC_COMM_SELF MPI_COMM_SELF
#define PETSC_ERR_POINTER 68
#define PETSC_ERR_FP 72
#define PETSC_ERR_SUP 56
#define PETSC_ERR_ARG_WRONG 62
#define CHKERRQ(ierr) do { if(ierr) return ierr; } while(0)
#define SETERRQ(comm, code, msg) return code

/// nlocal + nghost entries; the "local form" of a ghosted Vec is the Vec itself in a serial run
struct _p_Vec { std::vector<PetscScalar> a; PetscInt nlocal; PetscInt nghost; };
typedef _p_Vec* Vec;
typedef struct _p_PetscObject* PetscObject;

static inline PetscErrorCode VecGetArray(Vec v, PetscScalar **p) { *p = v->a.data(); return 0; }
static inline PetscErrorCode VecRestoreArray(Vec, PetscScalar **p) { *p = NULL; return 0; }
static inline PetscErrorCode VecGetArrayRead(Vec v, const PetscScalar **p) { *p = v->a.data(); return 0; }
static inline PetscErrorCode VecRestoreArrayRead(Vec, const PetscScalar **p) { *p = NULL; return 0; }
static inline PetscErrorCode VecGetLocalSize(Vec v, PetscInt *n) { *n = v->nlocal; return 0; }
static inline PetscErrorCode VecGhostGetLocalForm(Vec v, Vec *l) { *l = v; return 0; }
static inline PetscErrorCode VecGhostRestoreLocalForm(Vec, Vec *l) { *l = NULL; return 0; }
static inline PetscErrorCode VecGhostUpdateBegin(Vec, InsertMode, ScatterMode) { return 0; }
static inline PetscErrorCode VecGhostUpdateEnd(Vec, InsertMode, ScatterMode) { return 0; }
static inline PetscErrorCode VecDestroy(Vec *v) { delete *v; *v = NULL; return 0; }
static inline PetscErrorCode VecDuplicate(Vec v, Vec *w) { *w = new _p_Vec(*v); return 0; }
static inline PetscErrorCode VecSet(Vec v, PetscScalar s) { for(auto& x : v->a) x = s; return 0; }
typedef struct _p_PetscViewer* PetscViewer;
typedef enum { FILE_MODE_READ = 0, FILE_MODE_WRITE = 1 } PetscFileMode;
static inline PetscErrorCode PetscViewerBinaryOpen(MPI_Comm, const char*, PetscFileMode, PetscViewer *v) { *v = NULL; return PETSC_ERR_SUP; }
static inline PetscErrorCode PetscViewerDestroy(PetscViewer *v) { *v = NULL; return 0; }
static inline PetscErrorCode VecView(Vec, PetscViewer) { return PETSC_ERR_SUP; }
static inline PetscErrorCode PetscObjectGetComm(PetscObject, MPI_Comm *c) { *c = MPI_COMM_WORLD; return 0; }
#endif
