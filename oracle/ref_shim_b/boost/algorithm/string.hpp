/* boost::split / boost::is_any_of as the reference's mesh sources use them (file-extension and option splitting).
 * TEST INFRASTRUCTURE ONLY. */
#ifndef FVENS_B200_BOOST_STRING_LITE
#define FVENS_B200_BOOST_STRING_LITE
#include <string>
#include <vector>
#include <cctype>
namespace boost {
struct is_any_of_pred { std::string chars; bool operator()(const char c) const { return chars.find(c) != std::string::npos; } };
inline is_any_of_pred is_any_of(const std::string& chars) { return is_any_of_pred{chars}; }
/// token_compress_off semantics: adjacent separators produce empty tokens
template <typename Pred>
std::vector<std::string>& split(std::vector<std::string>& out, const std::string& s, Pred pred) {
	out.clear();
	std::string cur;
	for(const char c : s) { if(pred(c)) { out.push_back(cur); cur.clear(); } else cur += c; }
	out.push_back(cur);
	return out;
}
inline std::string to_upper_copy(std::string s) { for(char& c : s) c = (char)std::toupper((unsigned char)c); return s; }
inline std::string to_lower_copy(std::string s) { for(char& c : s) c = (char)std::tolower((unsigned char)c); return s; }
}
#endif
