// Stub of yaml-cpp for building the reference core (TEST INFRASTRUCTURE): only
// SettingsModel::LogAll(const YAML::Node&) refers to it (SettingsModel.cpp:1034-1045), off the path.
#ifndef HB_ORACLE_YAML_STUB_H
#define HB_ORACLE_YAML_STUB_H
#include <string>
namespace YAML {
struct Node {
    template <class T> T as() const { return T(); }
    Node operator[](const std::string&) const { return Node(); }
    explicit operator bool() const { return false; }
};
}  // namespace YAML
#endif
