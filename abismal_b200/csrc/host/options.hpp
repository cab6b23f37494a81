// Command-line parsing with the syntax of the reference's OptionParser
// (src/smithlab_cpp/OptionParser.cpp:226-270, :396-438): an option matches a
// token equal to "-x", "-long" or the bare "long"; bools take no value;
// every other option consumes the next token; duplicates throw; options are
// scanned in declaration order.
#ifndef ABISMAL_B200_OPTIONS_HPP
#define ABISMAL_B200_OPTIONS_HPP

#include <cstdint>
#include <string>
#include <vector>

namespace ab2 {

class Options {
public:
  void add(const std::string &long_name, char short_name, const std::string &descr, bool required,
           std::string &val);
  void add(const std::string &long_name, char short_name, const std::string &descr, bool required,
           uint32_t &val);
  void add(const std::string &long_name, char short_name, const std::string &descr, bool required, double &val);
  void add(const std::string &long_name, char short_name, const std::string &descr, bool required, bool &val);
  // argv[0] is skipped; returns leftover (positional) arguments. Throws std::runtime_error.
  std::vector<std::string> parse(int argc, char *const argv[]);
  bool option_missing() const { return !first_missing_.empty(); }
  std::string option_missing_message() const { return "option missing: " + first_missing_; }
  std::string help_message(const std::string &prog, const std::string &noflag) const;

private:
  enum Kind { kString, kUint, kDouble, kBool };
  struct Opt {
    std::string long_name;
    char short_name;
    std::string descr;
    bool required;
    Kind kind;
    void *target;
    bool specified = false;
  };
  bool match(const Opt &o, const std::string &tok) const;
  void assign(Opt &o, const std::string &val);
  std::vector<Opt> opts_;
  std::string first_missing_;
};

}  // namespace ab2
#endif
