#pragma once
struct sqlite3; struct sqlite3_stmt;
