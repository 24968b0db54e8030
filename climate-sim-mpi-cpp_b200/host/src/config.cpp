// config.cpp — YAML + CLI configuration with the reference's precedence and quirks.
//
// Behaviour restated from src/io.cpp:30-376:
//   - nested blocks grid/physics/time/output or the same keys flat at the top level (io.cpp:88-123,
//     143-148); a scalar `bc:` sets all four sides, a map sets the sides it names (io.cpp:125-141);
//   - ic.* is read from `ic:` directly (ic.mode, ic.preset, ic.A, …; io.cpp:150-168) — keys nested
//     one level deeper (dev.yaml's ic.params.*) are ignored, as upstream;
//   - CLI flags `--key=value` or `--key value`; unknown flags are ignored; CLI wins over YAML
//     (io.cpp:180-376); ic.var is parsed but never applied (io.cpp:305 vs 347-360);
//   - validate() runs after loading and after merging, with the reference's messages.
// The YAML reader covers what the reference's files and tests use: block maps by indentation, flow
// maps `{ k: v, … }`, plain/quoted scalars, comments.  It is not a general YAML parser.
#include <algorithm>
#include <cctype>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>

#include "csim_driver.hpp"

namespace {

std::string lower(std::string s) {
    std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return std::tolower(c); });
    return s;
}
std::string trim(const std::string& s) {
    size_t a = 0, b = s.size();
    while (a < b && std::isspace(static_cast<unsigned char>(s[a]))) ++a;
    while (b > a && std::isspace(static_cast<unsigned char>(s[b - 1]))) --b;
    return s.substr(a, b - a);
}
std::string unquote(const std::string& s) {
    if (s.size() >= 2 && ((s.front() == '"' && s.back() == '"') || (s.front() == '\'' && s.back() == '\'')))
        return s.substr(1, s.size() - 2);
    return s;
}
// strip a trailing comment that is not inside quotes
std::string strip_comment(const std::string& s) {
    char q = 0;
    for (size_t i = 0; i < s.size(); ++i) {
        const char c = s[i];
        if (q) {
            if (c == q) q = 0;
        } else if (c == '"' || c == '\'') {
            q = c;
        } else if (c == '#' && (i == 0 || std::isspace(static_cast<unsigned char>(s[i - 1])))) {
            return s.substr(0, i);
        }
    }
    return s;
}

// A node is a scalar or a map of nodes (all the reference's configs need).
struct Node {
    bool is_map = false;
    std::string scalar;
    std::map<std::string, Node> kids;
    const Node* get(const std::string& k) const {
        if (!is_map) return nullptr;
        auto it = kids.find(k);
        return it == kids.end() ? nullptr : &it->second;
    }
};

Node parse_flow_map(const std::string& text) {  // "{ a: 1, b: x }"
    Node n;
    n.is_map = true;
    const std::string inner = trim(text.substr(1, text.size() - 2));
    size_t pos = 0;
    while (pos < inner.size()) {
        size_t end = pos;
        char q = 0;
        int depth = 0;
        for (; end < inner.size(); ++end) {
            const char c = inner[end];
            if (q) {
                if (c == q) q = 0;
            } else if (c == '"' || c == '\'') {
                q = c;
            } else if (c == '{') {
                ++depth;
            } else if (c == '}') {
                --depth;
            } else if (c == ',' && depth == 0) {
                break;
            }
        }
        const std::string item = trim(inner.substr(pos, end - pos));
        pos = end + 1;
        if (item.empty()) continue;
        const size_t colon = item.find(':');
        if (colon == std::string::npos) throw std::runtime_error("yaml: expected key: value in flow map: " + item);
        const std::string key = unquote(trim(item.substr(0, colon)));
        const std::string val = trim(item.substr(colon + 1));
        if (!val.empty() && val.front() == '{' && val.back() == '}') {
            n.kids[key] = parse_flow_map(val);
        } else {
            Node s;
            s.scalar = unquote(val);
            n.kids[key] = s;
        }
    }
    return n;
}

struct Line {
    int indent;
    std::string key, value;
};

Node parse_block(const std::vector<Line>& lines, size_t& i, int indent) {
    Node n;
    n.is_map = true;
    while (i < lines.size() && lines[i].indent >= indent) {
        if (lines[i].indent > indent) throw std::runtime_error("yaml: unexpected indentation near " + lines[i].key);
        const Line& L = lines[i];
        ++i;
        if (!L.value.empty()) {
            if (L.value.front() == '{' && L.value.back() == '}') {
                n.kids[L.key] = parse_flow_map(L.value);
            } else {
                Node s;
                s.scalar = unquote(L.value);
                n.kids[L.key] = s;
            }
        } else if (i < lines.size() && lines[i].indent > indent) {
            n.kids[L.key] = parse_block(lines, i, lines[i].indent);
        } else {
            n.kids[L.key] = Node();  // empty scalar
        }
    }
    return n;
}

Node load_yaml(const std::string& path) {
    std::ifstream in(path);
    if (!in) throw std::runtime_error("bad file: " + path);  // yaml-cpp: YAML::BadFile
    std::vector<Line> lines;
    std::string raw;
    while (std::getline(in, raw)) {
        const std::string body = strip_comment(raw);
        if (trim(body).empty() || trim(body) == "---") continue;
        int indent = 0;
        while (indent < static_cast<int>(body.size()) && body[static_cast<size_t>(indent)] == ' ') ++indent;
        const std::string t = trim(body);
        const size_t colon = t.find(':');
        if (colon == std::string::npos) throw std::runtime_error("yaml: expected key: value: " + t);
        lines.push_back(Line{indent, unquote(trim(t.substr(0, colon))), trim(t.substr(colon + 1))});
    }
    size_t i = 0;
    return parse_block(lines, i, lines.empty() ? 0 : lines[0].indent);
}

// yaml-cpp's as<int>/as<double> reject trailing junk; so do these
int as_int(const Node& n) {
    size_t used = 0;
    const int v = std::stoi(n.scalar, &used);
    if (used != n.scalar.size()) throw std::runtime_error("yaml: bad conversion to int: " + n.scalar);
    return v;
}
double as_double(const Node& n) {
    size_t used = 0;
    const double v = std::stod(n.scalar, &used);
    if (used != n.scalar.size()) throw std::runtime_error("yaml: bad conversion to double: " + n.scalar);
    return v;
}
void assign_if(const Node& n, const char* key, int& x) {
    if (const Node* k = n.get(key)) x = as_int(*k);
}
void assign_if(const Node& n, const char* key, double& x) {
    if (const Node* k = n.get(key)) x = as_double(*k);
}
void assign_if(const Node& n, const char* key, std::string& x) {
    if (const Node* k = n.get(key)) x = k->scalar;
}

bool starts_with(const std::string& s, const std::string& p) { return s.rfind(p, 0) == 0; }

}  // namespace

BCType bc_from_string(const std::string& s) {
    const std::string t = lower(s);
    if (t == "dirichlet" || t == "fixed") return BCType::Dirichlet;
    if (t == "neumann" || t == "noflux" || t == "zero-flux") return BCType::Neumann;
    if (t == "periodic" || t == "period") return BCType::Periodic;
    throw std::runtime_error("Unknown BC type: " + s);
}

std::string bc_to_string(BCType bc) {
    switch (bc) {
        case BCType::Neumann: return "neumann";
        case BCType::Periodic: return "periodic";
        default: return "dirichlet";  // Dirichlet and anything out of range (io.cpp:55)
    }
}

void SimConfig::validate() const {
    if (nx <= 0 || ny <= 0) throw std::runtime_error("nx/ny must be > 0");
    if (dx <= 0 || dy <= 0) throw std::runtime_error("dx/dy must be > 0");
    if (dt <= 0) throw std::runtime_error("dt must be > 0");
    if (steps <= 0) throw std::runtime_error("steps must be > 0");
    if (out_every < 1) throw std::runtime_error("out_every must be >= 1");
}

SimConfig load_yaml_file(const std::string& path) {
    SimConfig cfg;
    const Node root = load_yaml(path);
    const Node* g = root.get("grid");
    const Node& grid = g ? *g : root;
    assign_if(grid, "nx", cfg.nx);
    assign_if(grid, "ny", cfg.ny);
    assign_if(grid, "dx", cfg.dx);
    assign_if(grid, "dy", cfg.dy);
    const Node* p = root.get("physics");
    const Node& phys = p ? *p : root;
    assign_if(phys, "D", cfg.D);
    assign_if(phys, "vx", cfg.vx);
    assign_if(phys, "vy", cfg.vy);
    const Node* t = root.get("time");
    const Node& tm = t ? *t : root;
    assign_if(tm, "dt", cfg.dt);
    assign_if(tm, "steps", cfg.steps);
    assign_if(tm, "out_every", cfg.out_every);
    if (const Node* b = root.get("bc")) {
        if (!b->is_map) {
            cfg.bc.left = cfg.bc.right = cfg.bc.bottom = cfg.bc.top = bc_from_string(b->scalar);
        } else {
            if (const Node* s = b->get("left")) cfg.bc.left = bc_from_string(s->scalar);
            if (const Node* s = b->get("right")) cfg.bc.right = bc_from_string(s->scalar);
            if (const Node* s = b->get("bottom")) cfg.bc.bottom = bc_from_string(s->scalar);
            if (const Node* s = b->get("top")) cfg.bc.top = bc_from_string(s->scalar);
        }
    }
    if (const Node* o = root.get("output"))
        assign_if(*o, "prefix", cfg.output_prefix);
    else
        assign_if(root, "output_prefix", cfg.output_prefix);
    if (const Node* ic = root.get("ic")) {
        assign_if(*ic, "mode", cfg.ic.mode);
        assign_if(*ic, "preset", cfg.ic.preset);
        assign_if(*ic, "A", cfg.ic.A);
        assign_if(*ic, "sigma_frac", cfg.ic.sigma_frac);
        assign_if(*ic, "xc_frac", cfg.ic.xc_frac);
        assign_if(*ic, "yc_frac", cfg.ic.yc_frac);
        assign_if(*ic, "path", cfg.ic.path);
        assign_if(*ic, "var", cfg.ic.var);
    }
    cfg.validate();
    return cfg;
}

CLIOverrides parse_cli_overrides(const std::vector<std::string>& args) {
    CLIOverrides o;
    // value of `--key=value` or of `--key value`; nullopt if this argument is not `key`
    auto value_of = [&](size_t i, const std::string& key) -> std::optional<std::string> {
        const std::string& a = args[i];
        if (starts_with(a, "--" + key + "=")) return a.substr(key.size() + 3);
        if (a == "--" + key && i + 1 < args.size()) return args[i + 1];
        return std::nullopt;
    };
    auto set_int = [&](size_t i, const char* k, std::optional<int>& dst) {
        if (auto v = value_of(i, k)) {
            dst = std::stoi(*v);
            return true;
        }
        return false;
    };
    auto set_dbl = [&](size_t i, const char* k, std::optional<double>& dst) {
        if (auto v = value_of(i, k)) {
            dst = std::stod(*v);
            return true;
        }
        return false;
    };
    auto set_str = [&](size_t i, const char* k, std::optional<std::string>& dst) {
        if (auto v = value_of(i, k)) {
            dst = *v;
            return true;
        }
        return false;
    };
    auto set_bc = [&](size_t i, const char* k, std::optional<BCType>& dst) {
        const std::string& a = args[i];
        const std::string key = k;
        if (!(starts_with(a, "--" + key + "=") || a == "--" + key)) return false;
        if (auto v = value_of(i, key))
            if (!v->empty()) dst = bc_from_string(*v);
        return true;
    };
    for (size_t i = 0; i < args.size(); ++i) {
        if (set_int(i, "nx", o.nx) || set_int(i, "ny", o.ny) || set_dbl(i, "dx", o.dx) || set_dbl(i, "dy", o.dy)) continue;
        if (set_dbl(i, "D", o.D) || set_dbl(i, "vx", o.vx) || set_dbl(i, "vy", o.vy)) continue;
        if (set_dbl(i, "dt", o.dt) || set_int(i, "steps", o.steps) || set_int(i, "out_every", o.out_every)) continue;
        if (set_bc(i, "bc.left", o.bc_left) || set_bc(i, "bc.right", o.bc_right) ||
            set_bc(i, "bc.bottom", o.bc_bottom) || set_bc(i, "bc.top", o.bc_top))
            continue;
        if (set_str(i, "output.prefix", o.output_prefix) || set_str(i, "output_prefix", o.output_prefix)) continue;
        if (set_str(i, "ic.mode", o.ic.mode) || set_str(i, "ic.preset", o.ic.preset)) continue;
        if (set_dbl(i, "ic.A", o.ic.A) || set_dbl(i, "ic.sigma_frac", o.ic.sigma_frac) ||
            set_dbl(i, "ic.xc_frac", o.ic.xc_frac) || set_dbl(i, "ic.yc_frac", o.ic.yc_frac))
            continue;
        if (set_str(i, "ic.path", o.ic.path) || set_str(i, "ic.var", o.ic.var)) continue;
        // anything else (e.g. --bc=periodic, --config=…) is silently ignored, as upstream
    }
    return o;
}

SimConfig merged_config(const std::optional<std::string>& yaml_path, const std::vector<std::string>& cli_args) {
    SimConfig cfg;
    if (yaml_path && !yaml_path->empty()) cfg = load_yaml_file(*yaml_path);
    const CLIOverrides o = parse_cli_overrides(cli_args);
    if (o.nx) cfg.nx = *o.nx;
    if (o.ny) cfg.ny = *o.ny;
    if (o.dx) cfg.dx = *o.dx;
    if (o.dy) cfg.dy = *o.dy;
    if (o.D) cfg.D = *o.D;
    if (o.vx) cfg.vx = *o.vx;
    if (o.vy) cfg.vy = *o.vy;
    if (o.dt) cfg.dt = *o.dt;
    if (o.steps) cfg.steps = *o.steps;
    if (o.out_every) cfg.out_every = *o.out_every;
    if (o.bc_left) cfg.bc.left = *o.bc_left;
    if (o.bc_right) cfg.bc.right = *o.bc_right;
    if (o.bc_bottom) cfg.bc.bottom = *o.bc_bottom;
    if (o.bc_top) cfg.bc.top = *o.bc_top;
    if (o.output_prefix) cfg.output_prefix = *o.output_prefix;
    if (o.ic.mode) cfg.ic.mode = *o.ic.mode;
    if (o.ic.preset) cfg.ic.preset = *o.ic.preset;
    if (o.ic.A) cfg.ic.A = *o.ic.A;
    if (o.ic.sigma_frac) cfg.ic.sigma_frac = *o.ic.sigma_frac;
    if (o.ic.xc_frac) cfg.ic.xc_frac = *o.ic.xc_frac;
    if (o.ic.yc_frac) cfg.ic.yc_frac = *o.ic.yc_frac;
    if (o.ic.path) cfg.ic.path = *o.ic.path;
    // o.ic.var is parsed and dropped (io.cpp:347-360 has no line for it)
    cfg.validate();
    return cfg;
}
