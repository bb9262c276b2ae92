// host/compat: the slice of boost::program_options that the TWILIGHT CLI uses. TEST INFRASTRUCTURE ONLY —
// exists so the unmodified reference host sources compile in an image without Boost headers.
//
// Supported: options_description(caption[, width]), add_options()("long,s"[, value<T>()[->default_value(v)]], "help"),
// add(), operator<<, variables_map::{count, operator[]}.as<T>(), command_line_parser(argc, argv).options(d).run(),
// store(), notify(). T in {std::string, int, float, double}. Accepts --long value, --long=value, -s value, -svalue.
#pragma once
// The real Boost headers pull these in transitively and the reference sources rely on it.
#include <algorithm>
#include <atomic>
#include <climits>
#include <cmath>
#include <functional>
#include <cstdlib>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <typeinfo>
#include <vector>

namespace boost {
namespace program_options {

class error : public std::runtime_error {
  public:
    explicit error(const std::string &what) : std::runtime_error(what) {}
};

struct value_semantic {
    virtual ~value_semantic() {}
    virtual bool has_default() const = 0;
    virtual std::shared_ptr<void> parse(const std::string &text) const = 0;
    virtual std::shared_ptr<void> default_holder() const = 0;
    virtual const std::type_info &type() const = 0;
    virtual std::string default_text() const = 0;
};

template <typename T>
class typed_value : public value_semantic {
  public:
    typed_value *default_value(const T &v) {
        default_ = std::make_shared<T>(v);
        return this;
    }
    bool has_default() const override { return static_cast<bool>(default_); }
    std::shared_ptr<void> parse(const std::string &text) const override {
        if constexpr (std::is_same<T, std::string>::value) {
            return std::make_shared<T>(text);
        } else {
            std::istringstream in(text);
            T v;
            in >> v;
            if (in.fail() || !in.eof()) throw error("the argument ('" + text + "') for an option is invalid");
            return std::make_shared<T>(v);
        }
    }
    std::shared_ptr<void> default_holder() const override { return default_; }
    const std::type_info &type() const override { return typeid(T); }
    std::string default_text() const override {
        if (!default_) return "";
        std::ostringstream out;
        out << *default_;
        return out.str();
    }

  private:
    std::shared_ptr<T> default_;
};

template <typename T>
typed_value<T> *value() {
    return new typed_value<T>();
}

struct option_entry {
    std::string longName;
    std::string shortName;
    std::string help;
    std::shared_ptr<value_semantic> semantic; // null => flag
};

class options_description;

class options_description_easy_init {
  public:
    explicit options_description_easy_init(options_description *owner) : owner_(owner) {}
    options_description_easy_init &operator()(const char *name, const char *help);
    options_description_easy_init &operator()(const char *name, value_semantic *s, const char *help = "");

  private:
    options_description *owner_;
};

class options_description {
  public:
    options_description(const std::string &caption = "", unsigned width = 80) : caption_(caption), width_(width) {}
    options_description_easy_init add_options() { return options_description_easy_init(this); }
    options_description &add(const options_description &other) {
        groups_.push_back(other);
        return *this;
    }
    void push(const option_entry &e) { entries_.push_back(e); }
    const option_entry *find_long(const std::string &n) const {
        for (auto &e : entries_) if (e.longName == n) return &e;
        for (auto &g : groups_) if (auto *p = g.find_long(n)) return p;
        return nullptr;
    }
    const option_entry *find_short(const std::string &n) const {
        for (auto &e : entries_) if (!e.shortName.empty() && e.shortName == n) return &e;
        for (auto &g : groups_) if (auto *p = g.find_short(n)) return p;
        return nullptr;
    }
    void collect(std::vector<const option_entry *> &out) const {
        for (auto &e : entries_) out.push_back(&e);
        for (auto &g : groups_) g.collect(out);
    }
    void print(std::ostream &os) const {
        if (!caption_.empty()) os << caption_ << ":\n";
        for (auto &e : entries_) {
            std::string left = "  ";
            if (!e.shortName.empty()) left += "-" + e.shortName + " [ --" + e.longName + " ]";
            else left += "--" + e.longName;
            if (e.semantic) {
                left += " arg";
                if (e.semantic->has_default()) left += " (=" + e.semantic->default_text() + ")";
            }
            os << left;
            if (left.size() < 40) os << std::string(40 - left.size(), ' ');
            else os << "\n" << std::string(40, ' ');
            os << e.help << "\n";
        }
        for (auto &g : groups_) {
            os << "\n";
            g.print(os);
        }
    }

  private:
    std::string caption_;
    unsigned width_;
    std::vector<option_entry> entries_;
    std::vector<options_description> groups_;
};

inline std::ostream &operator<<(std::ostream &os, const options_description &d) {
    d.print(os);
    return os;
}

inline options_description_easy_init &options_description_easy_init::operator()(const char *name, const char *help) {
    return (*this)(name, nullptr, help);
}
inline options_description_easy_init &options_description_easy_init::operator()(const char *name, value_semantic *s, const char *help) {
    option_entry e;
    std::string n(name);
    auto comma = n.find(',');
    if (comma == std::string::npos) e.longName = n;
    else {
        e.longName = n.substr(0, comma);
        e.shortName = n.substr(comma + 1);
    }
    e.help = help ? help : "";
    e.semantic.reset(s);
    owner_->push(e);
    return *this;
}

class variable_value {
  public:
    variable_value() {}
    variable_value(std::shared_ptr<void> v, const std::type_info *t) : v_(v), t_(t) {}
    template <typename T>
    const T &as() const {
        if (!v_ || *t_ != typeid(T)) throw error("bad option value cast");
        return *static_cast<const T *>(v_.get());
    }
    bool empty() const { return !v_; }

  private:
    std::shared_ptr<void> v_;
    const std::type_info *t_ = nullptr;
};

class variables_map {
  public:
    std::size_t count(const std::string &name) const { return values_.count(name); }
    const variable_value &operator[](const std::string &name) const {
        static const variable_value none;
        auto it = values_.find(name);
        return it == values_.end() ? none : it->second;
    }
    void set(const std::string &name, const variable_value &v) { values_[name] = v; }

  private:
    std::map<std::string, variable_value> values_;
};

struct parsed_options {
    const options_description *desc = nullptr;
    std::vector<std::pair<const option_entry *, std::string>> items; // value text ("" for flags)
};

class command_line_parser {
  public:
    command_line_parser(int argc, const char *const *argv) {
        for (int i = 1; i < argc; ++i) args_.push_back(argv[i]);
    }
    command_line_parser &options(const options_description &d) {
        desc_ = &d;
        return *this;
    }
    parsed_options run() {
        parsed_options out;
        out.desc = desc_;
        for (std::size_t i = 0; i < args_.size(); ++i) {
            const std::string &a = args_[i];
            const option_entry *e = nullptr;
            std::string inlineVal;
            bool hasInline = false;
            if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
                std::string body = a.substr(2);
                auto eq = body.find('=');
                if (eq != std::string::npos) {
                    inlineVal = body.substr(eq + 1);
                    body = body.substr(0, eq);
                    hasInline = true;
                }
                e = desc_->find_long(body);
                if (!e) throw error("unrecognised option '" + a + "'");
            } else if (a.size() >= 2 && a[0] == '-') {
                e = desc_->find_short(a.substr(1, 1));
                if (!e) throw error("unrecognised option '" + a + "'");
                if (a.size() > 2) {
                    inlineVal = a.substr(2);
                    hasInline = true;
                }
            } else {
                throw error("too many positional options have been specified on the command line");
            }
            if (e->semantic) {
                if (!hasInline) {
                    if (i + 1 >= args_.size()) throw error("the required argument for option '--" + e->longName + "' is missing");
                    inlineVal = args_[++i];
                }
                out.items.emplace_back(e, inlineVal);
            } else {
                out.items.emplace_back(e, "");
            }
        }
        return out;
    }

  private:
    std::vector<std::string> args_;
    const options_description *desc_ = nullptr;
};

inline void store(const parsed_options &parsed, variables_map &vm) {
    for (auto &it : parsed.items) {
        const option_entry *e = it.first;
        if (vm.count(e->longName)) continue;
        if (e->semantic) vm.set(e->longName, variable_value(e->semantic->parse(it.second), &e->semantic->type()));
        else vm.set(e->longName, variable_value(std::make_shared<bool>(true), &typeid(bool)));
    }
    std::vector<const option_entry *> all;
    parsed.desc->collect(all);
    for (auto *e : all) {
        if (e->semantic && e->semantic->has_default() && !vm.count(e->longName))
            vm.set(e->longName, variable_value(e->semantic->default_holder(), &e->semantic->type()));
    }
}

inline void notify(variables_map &) {}

} // namespace program_options
} // namespace boost
