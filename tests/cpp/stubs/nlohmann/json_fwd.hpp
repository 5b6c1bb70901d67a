#pragma once
#include <map>
#include <string>
#include <vector>
namespace nlohmann {
class json {
public:
    json(); template <class T> json(const T&);
    template <class T> json& operator=(const T&);
    json& operator[](const std::string&); const json& operator[](const std::string&) const; json& operator[](const char*); json& operator[](size_t);
    const json& at(const std::string&) const; json& at(const std::string&);
    template <class T> T get() const; template <class T> operator T() const;
    bool contains(const std::string&) const; size_t size() const; bool empty() const; bool is_null() const;
    void push_back(const json&);
    struct iterator { json& operator*(); iterator& operator++(); bool operator!=(const iterator&) const; const std::string& key() const; json& value(); };
    iterator begin(); iterator end(); iterator begin() const; iterator end() const;
    struct items_proxy { iterator begin(); iterator end(); };
    items_proxy items() const;
    static json array(); static json object();
};
}  // namespace nlohmann
