// placeholder until the model runner lands
#pragma once
