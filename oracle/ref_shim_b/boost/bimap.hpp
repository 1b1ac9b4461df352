/* The one use of boost::bimap on the residual path is the boundary-condition name table (spatial/abctypemap.cpp):
 * insert(value_type(l, r)), left.find(l)->second, right.at(r). TEST INFRASTRUCTURE ONLY. */
#ifndef FVENS_B200_BIMAP_LITE
#define FVENS_B200_BIMAP_LITE
#include <map>
namespace boost {
template <typename L, typename R>
class bimap {
public:
	struct value_type { L l; R r; value_type(const L& a, const R& b) : l(a), r(b) {} };
	std::map<L,R> left;
	std::map<R,L> right;
	void insert(const value_type& v) { left[v.l] = v.r; right[v.r] = v.l; }
};
}
#endif
