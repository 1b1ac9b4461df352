/* Names only: mesh/meshpartitioning.cpp calls Scotch in ScotchRGMPartitioner::compute_partition, which the reference
 * itself has commented out of its mesh set-up (ameshutils.cpp:122-123) and the harness never calls. TEST INFRASTRUCTURE ONLY. */
#ifndef FVENS_B200_SCOTCH_LITE
#define FVENS_B200_SCOTCH_LITE
typedef int SCOTCH_Num;
typedef struct { int dummy; } SCOTCH_Graph;
typedef struct { int dummy; } SCOTCH_Strat;
static inline SCOTCH_Graph* SCOTCH_graphAlloc() { return new SCOTCH_Graph; }
static inline SCOTCH_Strat* SCOTCH_stratAlloc() { return new SCOTCH_Strat; }
static inline int SCOTCH_graphBuild(SCOTCH_Graph*, SCOTCH_Num, SCOTCH_Num, const SCOTCH_Num*, const SCOTCH_Num*, const SCOTCH_Num*,
                                    const SCOTCH_Num*, SCOTCH_Num, const SCOTCH_Num*, const SCOTCH_Num*) { return 1; }
static inline int SCOTCH_graphCheck(const SCOTCH_Graph*) { return 1; }
static inline int SCOTCH_stratInit(SCOTCH_Strat*) { return 1; }
static inline int SCOTCH_graphPart(SCOTCH_Graph*, SCOTCH_Num, SCOTCH_Strat*, SCOTCH_Num*) { return 1; }
static inline void SCOTCH_graphExit(SCOTCH_Graph*) {}
static inline void SCOTCH_stratExit(SCOTCH_Strat*) {}
static inline void SCOTCH_memFree(void *p) { (void)p; }
#endif
