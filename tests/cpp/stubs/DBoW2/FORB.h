#pragma once
namespace DBoW2 { class FORB {}; }
