#include <string>
