/* Names only: utilities/aoptionparser.hpp declares a function over these Boost types; nothing on the explicit path
 * calls it. TEST INFRASTRUCTURE ONLY. */
#ifndef FVENS_B200_PO_LITE
#define FVENS_B200_PO_LITE
#include <map>
#include <string>
namespace boost { namespace program_options {
class options_description { public: options_description() {} explicit options_description(const std::string&) {} };
class variables_map : public std::map<std::string,std::string> {};
}}
#endif
