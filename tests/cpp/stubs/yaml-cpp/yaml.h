#pragma once
#include <string>
#include <vector>
namespace YAML {
class Node {
public:
    Node(); template <class T> Node(const T&);
    template <class T> T as() const; template <class T, class D> T as(const D&) const;
    Node operator[](const std::string&) const; Node operator[](const char*) const; Node operator[](size_t) const;
    template <class T> Node& operator=(const T&);
    explicit operator bool() const; bool operator!() const; bool IsDefined() const; bool IsSequence() const; bool IsMap() const; size_t size() const;
    struct iterator { Node operator*() const; iterator& operator++(); bool operator!=(const iterator&) const; };
    iterator begin() const; iterator end() const;
};
Node LoadFile(const std::string&); Node Load(const std::string&); Node Clone(const Node&);
}  // namespace YAML
