#include <petscvec.h>
