// .essa world files -> World (the step before the hot path: "same .essa inputs" made literal).
//
// Grammar and defaults follow the reference's src/ConfigLoader.cpp:17-30 (statement list),
// :110-185 (statement kinds), :187-202 (key=value), :61-78 (distances are parsed as *float* and
// carry a unit suffix _m / _km / _AU), :204-270 (application order = file order).
//
//   planet           mass radius colorr colorg colorb name trail_length posx posy posz velx vely velz
//   orbiting_planet  ... around direction orbit_position orbit_tilt  (apoapsis periapsis | major_axis eccentrity)
//   light_source     <Name>
//
// Two passes here (scan into records, then build objects) instead of the reference's
// parser-combinator + variant visitor; the observable result is the same object list.
#include <cctype>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>

#include "facade.hpp"

namespace essa {

namespace {

struct Cursor {
    std::string const& text;
    size_t at = 0;
    bool done() const { return at >= text.size(); }
    int look() const { return done() ? -1 : static_cast<unsigned char>(text[at]); }
    void blanks() {
        while (!done() && std::isspace(look()))
            at++;
    }
    template<class Pred> std::string take(Pred keep) {
        size_t from = at;
        while (!done() && keep(look()))
            at++;
        return text.substr(from, at - from);
    }
    std::string position() const {
        size_t line = 1, col = 1;
        for (size_t i = 0; i < at && i < text.size(); i++) {
            if (text[i] == '\n') {
                line++;
                col = 1;
            } else
                col++;
        }
        return std::to_string(line) + ":" + std::to_string(col);
    }
};

struct Failure {
    std::string what;
};

struct Record {
    std::string kind;
    std::string bare; // light_source's operand
    std::map<std::string, std::string> kv;

    bool has(char const* k) const { return kv.count(k) != 0; }
    std::string text(char const* k) const {
        auto it = kv.find(k);
        return it == kv.end() ? std::string() : it->second;
    }
    double number(char const* k, double fallback = 0) const {
        auto it = kv.find(k);
        if (it == kv.end())
            return fallback;
        char* end = nullptr;
        double v = std::strtod(it->second.c_str(), &end);
        if (end == it->second.c_str())
            throw Failure { "Invalid number" };
        return v;
    }
    int integer(char const* k, int fallback) const {
        auto it = kv.find(k);
        if (it == kv.end())
            return fallback;
        char* end = nullptr;
        long v = std::strtol(it->second.c_str(), &end, 10);
        if (end == it->second.c_str())
            throw Failure { std::string(k) + " is an invalid number" };
        return static_cast<int>(v);
    }
    uint8_t byte(char const* k, uint8_t fallback) const {
        auto it = kv.find(k);
        if (it == kv.end())
            return fallback;
        int v = integer(k, fallback);
        if (v < 0 || v > 255)
            throw Failure { std::string(k) + " must be in range 0..256" };
        return static_cast<uint8_t>(v);
    }
    // Distance holds a float; "<number>_<unit>" or a bare number of metres
    float distance(char const* k) const {
        auto it = kv.find(k);
        if (it == kv.end())
            return 0.0f;
        size_t us = it->second.find('_');
        if (us == std::string::npos)
            return static_cast<float>(number(k));
        std::string const digits = it->second.substr(0, us), unit = it->second.substr(us + 1);
        char* end = nullptr;
        float v = std::strtof(digits.c_str(), &end);
        if (end == digits.c_str())
            throw Failure { "Invalid number" };
        if (unit == "m")
            return v;
        if (unit == "km")
            return v * 1000;
        if (unit == "AU")
            return static_cast<float>(v * ESSA_AU);
        throw Failure { "Invalid unit" };
    }
};

bool word_char(int c) { return std::isalpha(c) || c == '_'; }

std::vector<Record> scan(std::string const& text) {
    std::vector<Record> out;
    Cursor c { text };
    c.blanks();
    while (!c.done()) {
        Record r;
        r.kind = c.take(word_char);
        if (r.kind.empty())
            throw Failure { "Expected keyword at " + c.position() };
        c.blanks();
        if (r.kind == "planet" || r.kind == "orbiting_planet") {
            for (;;) {
                c.blanks();
                std::string key = c.take(word_char);
                c.blanks();
                if (c.look() != '=')
                    throw Failure { "Expected '=' at " + c.position() };
                c.at++;
                c.blanks();
                std::string value = c.take([](int ch) { return ch != ';' && !std::isspace(ch); });
                r.kv.emplace(std::move(key), std::move(value)); // first occurrence wins, as std::map::insert
                c.blanks();
                if (c.look() < 0)
                    throw Failure { "Expected ';' after property list at " + c.position() };
                if (!std::isalpha(c.look()))
                    break;
            }
        } else if (r.kind == "light_source") {
            r.bare = c.take([](int ch) { return std::isalnum(ch) != 0; });
        } else {
            throw Failure { "Invalid statement keyword: '" + r.kind + "' at " + c.position() };
        }
        c.blanks();
        if (c.look() != ';')
            throw Failure { "Expected ';' at " + c.position() };
        c.at++;
        c.blanks();
        out.push_back(std::move(r));
    }
    return out;
}

double degrees(double d) { return d * M_PI / 180.0; } // Util::Angle::degrees

void build(Record const& r, World& world) {
    if (r.kind == "light_source") {
        if (!world.set_light_source(r.bare))
            throw Failure { "Invalid planet for light source: '" + r.bare + "'" };
        return;
    }
    double const mass = r.number("mass");
    float const radius = r.distance("radius");
    Color const color { r.byte("colorr", 255), r.byte("colorg", 255), r.byte("colorb", 255), 255 };
    std::string const name = r.text("name");
    double const trail_length = r.integer("trail_length", 600);
    if (r.kind == "planet") {
        Vec3 pos { r.number("posx"), r.number("posy"), r.number("posz") };
        Vec3 vel { r.number("velx"), r.number("vely"), r.number("velz") };
        world.add_object(std::make_unique<Object>(mass, radius, pos, vel, color, name, static_cast<unsigned>(trail_length)));
        return;
    }
    bool const by_ap_pe = r.has("apoapsis") && r.has("periapsis");
    bool const by_maj_ecc = r.has("major_axis") && r.has("eccentrity");
    if (!by_ap_pe && !by_maj_ecc)
        throw Failure { "orbiting_planet must define apoapsis+periapsis or major_axis+eccentrity" };
    double const theta = degrees(r.number("orbit_position")), alpha = degrees(r.number("orbit_tilt"));
    bool const clockwise = r.text("direction") != "left";
    std::string const around_name = r.text("around");
    Object* around = world.get_object_by_name(around_name);
    if (!around)
        throw Failure { "Invalid planet for orbiting around: '" + around_name + "'" };
    if (by_ap_pe)
        world.add_object(around->create_object_relative_to_ap_pe(mass, radius, r.distance("apoapsis"), r.distance("periapsis"), clockwise,
            theta, alpha, color, name, 0.0));
    else
        world.add_object(around->create_object_relative_to_maj_ecc(mass, radius, r.distance("major_axis"), r.number("eccentrity"),
            clockwise, theta, alpha, color, name, 0.0));
}

} // namespace

std::string load_essa_text(std::string const& text, World& world) {
    try {
        for (auto const& rec : scan(text))
            build(rec, world);
    } catch (Failure const& f) {
        return f.what;
    }
    return {};
}

std::string load_essa_file(std::string const& filename, World& world) {
    std::ifstream in(filename);
    if (!in)
        return "open: cannot read " + filename;
    std::stringstream ss;
    ss << in.rdbuf();
    return load_essa_text(ss.str(), world);
}

} // namespace essa
