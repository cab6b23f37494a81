#include "options.hpp"

#include <sstream>
#include <stdexcept>

namespace ab2 {

void Options::add(const std::string &l, char s, const std::string &d, bool r, std::string &v) {
  opts_.push_back(Opt{l, s, d, r, kString, &v});
}
void Options::add(const std::string &l, char s, const std::string &d, bool r, uint32_t &v) {
  opts_.push_back(Opt{l, s, d, r, kUint, &v});
}
void Options::add(const std::string &l, char s, const std::string &d, bool r, double &v) {
  opts_.push_back(Opt{l, s, d, r, kDouble, &v});
}
void Options::add(const std::string &l, char s, const std::string &d, bool r, bool &v) {
  opts_.push_back(Opt{l, s, d, r, kBool, &v});
}

bool Options::match(const Opt &o, const std::string &tok) const {
  return o.long_name == tok ||
         (tok.size() > 1 && tok[0] == '-' &&
          (tok.substr(1) == o.long_name || (tok[1] == o.short_name && tok.size() == 2)));
}

void Options::assign(Opt &o, const std::string &val) {
  std::istringstream ss(val);
  switch (o.kind) {
    case kString: *static_cast<std::string *>(o.target) = val; break;
    case kBool: *static_cast<bool *>(o.target) = !*static_cast<bool *>(o.target); break;
    case kUint:
      if (!(ss >> *static_cast<uint32_t *>(o.target)))
        throw std::runtime_error("Invalid argument [" + val + "] to option [-" + o.long_name + "]");
      break;
    case kDouble:
      if (!(ss >> *static_cast<double *>(o.target)))
        throw std::runtime_error("Invalid argument [" + val + "] to option [-" + o.long_name + "]");
      break;
  }
}

std::vector<std::string> Options::parse(int argc, char *const argv[]) {
  std::vector<std::string> args(argv + 1, argv + argc);
  static const std::string dummy;
  for (Opt &o : opts_) {
    for (size_t i = 0; i < args.size();) {
      if (match(o, args[i])) {
        if (o.specified) throw std::runtime_error("duplicate use of option: " + o.long_name);
        if (i + 1 < args.size()) assign(o, args[i + 1]);
        else assign(o, dummy);
        o.specified = true;
        args.erase(args.begin() + static_cast<long>(i));
        if (o.kind != kBool && i < args.size()) args.erase(args.begin() + static_cast<long>(i));
      }
      else ++i;
    }
    if (!o.specified && o.required && first_missing_.empty()) {
      first_missing_ = o.short_name ? std::string("-") + o.short_name + ", -" + o.long_name
                                    : "    -" + o.long_name;
    }
  }
  return args;
}

std::string Options::help_message(const std::string &prog, const std::string &noflag) const {
  std::ostringstream ss;
  ss << "Usage: " << prog << " [OPTIONS] " << noflag << "\n\nOptions:\n";
  for (const Opt &o : opts_) {
    ss << "  ";
    if (o.short_name) ss << '-' << o.short_name << ", ";
    else ss << "    ";
    ss << '-' << o.long_name << "  " << o.descr << (o.required ? " [required]" : "") << '\n';
  }
  return ss.str();
}

}  // namespace ab2
