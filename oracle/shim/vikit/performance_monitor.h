#pragma once
namespace vk { class PerformanceMonitor {}; }
