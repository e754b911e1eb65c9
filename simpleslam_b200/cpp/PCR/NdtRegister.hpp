// NdtRegister lives in PointCloudRegister.hpp (kept as a separate header so the reference's `#include <PCR/NdtRegister.hpp>` lines keep working)
#pragma once
#include <PCR/PointCloudRegister.hpp>
