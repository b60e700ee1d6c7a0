/*
 * oracle/hhmarl_oracle.c -- TEST INFRASTRUCTURE (CPU oracle). Not part of the product path.
 *
 * Scalar C restatement of the reference's low-level environment step.  Each function cites
 * the reference lines it follows (paths relative to /root/reference).  The structure is the
 * reference's own (unit registry, id-ordered tick loop, event list) -- deliberately NOT the
 * lane-parallel structure of the CUDA kernels, so the two can check one another.
 *
 * Build: gcc -O2 -ffp-contract=off (CPython never fuses a*b+c).
 */
#include "hhmarl_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "geodesic.h"
#include "philox.h"

#define KIND_AC1 1
#define KIND_AC2 2
#define KIND_ROCKET 3

#define KNOTS_TO_MS 0.514444 /* cmano_simulator.py:21 */
#define TICK_SECS 1          /* cmano_simulator.py:80 */

#define OBS_AC1 26
#define OBS_AC2 24
#define OBS_ESC_AC1 30
#define OBS_ESC_AC2 29

/* rocket_unit.py:16-21: scipy interp1d(kind='quadratic') through (0,500),(10,2000),(20,1400),
 * (30,600) evaluated at t = 0..10 s (oracle/gen_rocket_table.py prints these from scipy). */
static const double ROCKET_SPEED[11] = {
    0x1.f400000000000p+8,  0x1.7b5ffffffffffp+9,  0x1.f0aaaaaaaaaacp+9,  0x1.2cf0000000000p+10,
    0x1.5b80000000000p+10, 0x1.8405555555554p+10, 0x1.a680000000001p+10, 0x1.c2f0000000000p+10,
    0x1.d955555555556p+10, 0x1.e9b0000000000p+10, 0x1.f400000000000p+10,
};

typedef struct {
  int active; /* key present in sim.active_units */
  int kind;
  int id;
  double lat, lon, heading, speed;
  double new_heading, new_speed;
  /* aircraft */
  double max_speed;
  double cannon_remain_secs, cannon_current_burst_secs, cannon_max;
  int missile_remain, rocket_max;
  int actual_missile; /* unit id of the Rocket object, 0 = None */
  int group;          /* 0 "agent", 1 "opp" */
  int ac_type;
  int friendly_check;
  /* rocket */
  int target, source;
  long firing_time;
} unit_t;

typedef struct {
  int origin, killer, destroyed;
} event_t;

struct orc_env {
  orc_args_t args;
  int total_num;
  /* CmanoSimulator (cmano_simulator.py:79-93) */
  unit_t units[ORC_MAX_UNITS + 1];
  int next_unit_id;
  long utc_time;
  /* HHMARLBaseEnv (env_base.py:38-53) */
  int steps, alive_agents, alive_opps;
  int opp_to_attack[ORC_MAX_AC + 1];
  int missile_wait[ORC_MAX_AC + 1];
  int hardcoded_opps_escaping, opps_escaping_time;
  int opp_mode;   /* 0 fight, 1 escape (env_hetero.py:23,59) */
  int policy_set; /* k of env_hetero.py:57 */
  orc_rng_t rng_g, rng_c;
  orc_policy_fn policy_fn;
  void* policy_user;
  int error; /* set where the reference would raise */
  /* HighLevelEnv (env_hier.py): opp_to_attack[i] is a LIST [[id, d_norm, d_raw], ...]; commander actions */
  struct { int id; double d_norm, d_raw; } ota_list[ORC_MAX_AC + 1][ORC_MAX_AC];
  int ota_n[ORC_MAX_AC + 1];
  int commander_actions[ORC_MAX_AC + 1]; /* -1 = None */
};

/* ------------------------------------------------------------------ small helpers */
/* CPython float.__mod__ (Objects/floatobject.c float_rem) */
static double pymod(double x, double m) {
  double r = fmod(x, m);
  if (r != 0) {
    if ((m < 0) != (r < 0)) r += m;
  } else {
    r = copysign(0.0, m);
  }
  return r;
}

static double clip(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* builtin round(x, 3): correctly rounded decimal, then nearest double */
static double pyround3(double x) {
  char buf[64];
  snprintf(buf, sizeof buf, "%.3f", x);
  return strtod(buf, 0);
}

/* angles.py:22-29 */
static double signed_heading_diff(double actual, double desired) {
  double delta = desired - actual;
  if (delta < -180) delta = 360 + delta;
  if (delta > 180) delta = -360 + delta;
  return delta;
}

/* angles.py:10-19 */
static double normalize_angle(double a) {
  while (a >= 360.0) a -= 360;
  while (a < 0.0) a += 360;
  return a;
}
static double sum_angles(double a, double b) { return normalize_angle(a + b); }

/* env_base.py:458-462 */
static double shifted_range(double x, double a, double b, double c, double d) {
  return c + ((d - c) / (b - a)) * (x - a);
}

static unit_t* U(orc_env_t* e, int id) { return &e->units[id]; }
static int unit_exists(const orc_env_t* e, int id) {
  return id >= 1 && id <= ORC_MAX_UNITS && e->units[id].active;
}

/* cmano_simulator.py:104-108 */
static int add_unit(orc_env_t* e, const unit_t* u) {
  int id = e->next_unit_id;
  if (id > ORC_MAX_UNITS) {
    e->error |= 4;
    return 0;
  }
  e->units[id] = *u;
  e->units[id].id = id;
  e->units[id].active = 1;
  e->next_unit_id += 1;
  return id;
}
/* cmano_simulator.py:110-114 */
static void remove_unit(orc_env_t* e, int id) {
  if (!e->units[id].active) e->error |= 8; /* KeyError in the reference */
  e->units[id].active = 0;
}

/* cmano_simulator.py:167-174 */
static double units_distance_km(const unit_t* a, const unit_t* b) {
  return orc_geodetic_distance_km(a->lat, a->lon, b->lat, b->lon);
}
static double units_bearing(const unit_t* from, const unit_t* to) {
  return orc_geodetic_bearing_deg(from->lat, from->lon, to->lat, to->lon);
}

/* map_limits.py:37-40,47-48 with MapLimits(7.0, 5.0, 7.0+ms, 5.0+ms) (env_base.py:43) */
static void relative_position(const orc_env_t* e, double lat, double lon, double* lat_rel,
                              double* lon_rel) {
  double left = 7.0, bottom = 5.0, right = 7.0 + e->args.map_size, top = 5.0 + e->args.map_size;
  *lat_rel = clip((lat - bottom) / (top - bottom), 0, 1);
  *lon_rel = clip((lon - left) / (right - left), 0, 1);
}
static int in_boundary(const orc_env_t* e, double lat, double lon) {
  double left = 7.0, bottom = 5.0, right = 7.0 + e->args.map_size, top = 5.0 + e->args.map_size;
  return left <= lon && lon <= right && bottom <= lat && lat <= top;
}

/* ------------------------------------------------------------------ unit setters */
/* ac1.py:58-67, ac2.py:54-63, rocket_unit.py:32-35 */
static void set_heading(orc_env_t* e, unit_t* u, double h) {
  if (h >= 360 || h < 0) e->error |= 1; /* the reference raises */
  u->new_heading = h;
}
static void set_speed(orc_env_t* e, unit_t* u, double s) {
  if (s > u->max_speed || s < 0) e->error |= 2;
  u->new_speed = s;
}
/* ac1.py:69-70, ac2.py:65-66 */
static void fire_cannon(unit_t* u) {
  double burst = u->ac_type == 1 ? 5 : 3;
  u->cannon_current_burst_secs = u->cannon_remain_secs < burst ? u->cannon_remain_secs : burst;
}
/* ac1.py:144-146 */
static int angle_in_radar_range(const unit_t* u, double angle) {
  double delta = fabs(signed_heading_diff(sum_angles(u->heading, 120 / 2.0), angle));
  return (int)delta <= (int)(120 / 2.0);
}
/* ac1.py:72-79 */
static void fire_missile(orc_env_t* e, unit_t* self, unit_t* opp_unit) {
  if (!self->actual_missile && self->missile_remain > 0) {
    double unit_distance = units_distance_km(self, opp_unit);
    if (unit_distance <= 111 && angle_in_radar_range(self, units_bearing(self, opp_unit))) {
      unit_t m; /* rocket_unit.py:23-30 */
      memset(&m, 0, sizeof m);
      m.kind = KIND_ROCKET;
      m.lat = self->lat;
      m.lon = self->lon;
      m.heading = self->heading;
      if (m.heading >= 360 || m.heading < 0) e->error |= 1; /* cmano_simulator.py:57-58 */
      m.speed = ROCKET_SPEED[0];
      m.new_heading = m.heading;
      m.firing_time = e->utc_time;
      m.target = opp_unit->id;
      m.source = self->id;
      m.friendly_check = self->friendly_check;
      self->actual_missile = add_unit(e, &m);
      self->missile_remain = self->missile_remain - 1 > 0 ? self->missile_remain - 1 : 0;
    }
  }
}

/* cmano_simulator.py:65-72 */
static void unit_base_update(unit_t* u) {
  if (u->speed > 0) {
    double lat2, lon2;
    orc_geodetic_direct(u->lat, u->lon, u->heading, u->speed * KNOTS_TO_MS * TICK_SECS, &lat2, &lon2);
    u->lat = lat2;
    u->lon = lon2;
  }
}

/* ac1.py:135-142, ac2.py:109-116 */
static int unit_in_cannon_range(const unit_t* self, const unit_t* u) {
  double range = self->ac_type == 1 ? 2.0 : 4.5;
  double width = self->ac_type == 1 ? 10 : 7;
  double distance = units_distance_km(self, u);
  if (distance < range) {
    double bearing = units_bearing(self, u);
    double delta = fabs(signed_heading_diff(self->heading, bearing));
    return delta <= width / 2.0;
  }
  return 0;
}

/* ac1.py:81-133 (Rafale.update) and ac2.py:68-107 (RafaleLong.update) */
static void aircraft_update(orc_env_t* e, unit_t* self, event_t* events, int* n_events) {
  double max_deg_sec = self->ac_type == 1 ? 5 : 3.5;
  double max_knots_sec = self->ac_type == 1 ? 35 : 28;
  double hit_prob = self->ac_type == 1 ? 0.75 : 0.9;
  double burst_time = self->ac_type == 1 ? 5 : 3;
  int id;
  if (self->heading != self->new_heading) {
    double delta = signed_heading_diff(self->heading, self->new_heading);
    double max_deg = max_deg_sec * TICK_SECS;
    if (fabs(delta) <= max_deg) {
      self->heading = self->new_heading;
    } else {
      self->heading += delta >= 0 ? max_deg : -max_deg;
      self->heading = pymod(self->heading, 360);
    }
  }
  if (self->speed != self->new_speed) {
    double delta = self->new_speed - self->speed;
    double max_delta = max_knots_sec * TICK_SECS;
    if (fabs(delta) <= max_delta)
      self->speed = self->new_speed;
    else
      self->speed += delta >= 0 ? max_delta : -max_delta;
  }
  if (self->cannon_current_burst_secs > 0) {
    int snapshot[ORC_MAX_UNITS + 1], n_snap = 0, k;
    self->cannon_current_burst_secs = fmax(self->cannon_current_burst_secs - TICK_SECS, 0.0);
    self->cannon_remain_secs = fmax(self->cannon_remain_secs - TICK_SECS, 0.0);
    for (id = 1; id < e->next_unit_id; ++id)
      if (e->units[id].active) snapshot[n_snap++] = id;
    for (k = 0; k < n_snap; ++k) {
      unit_t* unit = U(e, snapshot[k]);
      if (unit->id != self->id) {
        if (unit->id <= e->args.num_agents + e->args.num_opps) {
          if (self->friendly_check || (self->group == 0 && unit->id >= e->args.num_agents + 1) ||
              (self->group == 1 && unit->id <= e->args.num_agents)) {
            if (unit->kind == KIND_AC1 || unit->kind == KIND_AC2) {
              if (unit_in_cannon_range(self, unit)) {
                if (orc_rng_random(&e->rng_c) < (hit_prob / (burst_time / TICK_SECS))) {
                  remove_unit(e, unit->id);
                  events[*n_events].origin = self->id;
                  events[*n_events].killer = self->id;
                  events[*n_events].destroyed = unit->id;
                  *n_events += 1;
                }
              }
            }
          }
        }
      }
    }
  }
  if (self->ac_type == 1 && self->actual_missile) { /* ac1.py:117-128 */
    if (!unit_exists(e, self->actual_missile)) {
      self->actual_missile = 0;
    } else {
      unit_t* m = U(e, self->actual_missile);
      double heading = clip(m->heading * orc_rng_uniform(&e->rng_g, 0.95, 1.05), 0, 359);
      set_heading(e, m, heading);
    }
  }
  unit_base_update(self);
}

/* rocket_unit.py:37-73 */
static void rocket_update(orc_env_t* e, unit_t* self, event_t* events, int* n_events) {
  long life_time;
  if (units_distance_km(self, U(e, self->target)) < 1 && unit_exists(e, self->target)) {
    remove_unit(e, self->id);
    remove_unit(e, self->target);
    events[*n_events].origin = self->id;
    events[*n_events].killer = self->source;
    events[*n_events].destroyed = self->target;
    *n_events += 1;
    return;
  }
  if (self->friendly_check) {
    int friendly_id = self->source == 2 ? 1 : 2;
    if (unit_exists(e, friendly_id)) {
      if (units_distance_km(self, U(e, friendly_id)) < 1) {
        remove_unit(e, self->id);
        remove_unit(e, friendly_id);
        events[*n_events].origin = self->id;
        events[*n_events].killer = self->source;
        events[*n_events].destroyed = friendly_id;
        *n_events += 1;
        return;
      }
    }
  }
  life_time = e->utc_time - self->firing_time;
  if (life_time > 10) {
    remove_unit(e, self->id);
    return;
  }
  if (self->heading != self->new_heading) {
    double delta = signed_heading_diff(self->heading, self->new_heading);
    double max_deg = 10 * TICK_SECS;
    if (fabs(delta) <= max_deg)
      self->heading = self->new_heading;
    else
      self->heading += delta >= 0 ? max_deg : -max_deg;
  }
  self->speed = ROCKET_SPEED[life_time];
  unit_base_update(self);
}

/* cmano_simulator.py:138-157 */
static int do_tick(orc_env_t* e, event_t* events) {
  int snapshot[ORC_MAX_UNITS + 1], n_snap = 0, k, id, n_events = 0;
  for (id = 1; id < e->next_unit_id; ++id)
    if (e->units[id].active) snapshot[n_snap++] = id;
  for (k = 0; k < n_snap; ++k) {
    unit_t* u = U(e, snapshot[k]);
    if (u->kind == KIND_ROCKET)
      rocket_update(e, u, events, &n_events);
    else
      aircraft_update(e, u, events, &n_events);
  }
  e->utc_time += TICK_SECS;
  return n_events;
}

/* ------------------------------------------------------------------ geometry features */
static void heading_vec(double heading, double* c, double* s) {
  double th = pymod(90 - heading, 360) * (M_PI / 180);
  *c = cos(th);
  *s = sin(th);
}

/* env_base.py:424-432 */
static double focus_angle(orc_env_t* e, int agent_id, int opp_id, int norm) {
  const unit_t* a = U(e, agent_id);
  const unit_t* o = U(e, opp_id);
  double u0, u1, v0, v1, x, deg;
  heading_vec(a->heading, &u0, &u1);
  v0 = o->lon - a->lon;
  v1 = o->lat - a->lat;
  x = clip((u0 * v0 + u1 * v1) / (sqrt(u0 * u0 + u1 * u1) * sqrt(v0 * v0 + v1 * v1) + 1e-10), -1, 1);
  deg = acos(x) * (180 / M_PI);
  return norm ? clip(deg / 180, 0, 1) : deg;
}

/* env_base.py:434-439 */
static double distance(orc_env_t* e, int agent_id, int opp_id, int norm) {
  const unit_t* a = U(e, agent_id);
  const unit_t* o = U(e, opp_id);
  double d = hypot(o->lon - a->lon, o->lat - a->lat);
  double ms = e->args.map_size;
  return norm ? shifted_range(d, 0, sqrt(2 * (ms * ms)), 0, 1) : d;
}

/* env_base.py:441-446 (norm=True only on the hot path) */
static double aspect_angle(orc_env_t* e, int agent_id, int opp_id) {
  double focus = focus_angle(e, agent_id, opp_id, 0);
  return clip((180 - focus) / 180, 0, 1);
}

/* env_base.py:448-456 (norm=True only) */
static double heading_diff(orc_env_t* e, int agent_id, int opp_id) {
  double u0, u1, w0, w1, x;
  heading_vec(U(e, agent_id)->heading, &u0, &u1);
  heading_vec(U(e, opp_id)->heading, &w0, &w1);
  x = clip((u0 * w0 + u1 * w1) / (sqrt(u0 * u0 + u1 * u1) * sqrt(w0 * w0 + w1 * w1) + 1e-10), -1, 1);
  return clip((acos(x) * (180 / M_PI)) / 180, 0, 1);
}

typedef struct {
  int id;
  double d_norm, d_raw;
} near_t;

/* env_base.py:400-422; returns count, list sorted by d_norm (stable) */
static int nearby_object(orc_env_t* e, int agent_id, int friendly, near_t* order) {
  int n = 0, i, j, start, end;
  if (friendly) {
    if (agent_id <= e->args.num_agents) {
      start = 1;
      end = e->args.num_agents + 1;
    } else {
      start = e->args.num_agents + 1;
      end = e->total_num + 1;
    }
    for (i = start; i < end; ++i) {
      if (i == agent_id) continue;
      if (unit_exists(e, i)) {
        order[n].id = i;
        order[n].d_norm = distance(e, agent_id, i, 1);
        order[n].d_raw = 0;
        ++n;
      }
    }
  } else {
    if (agent_id <= e->args.num_agents) {
      start = e->args.num_agents + 1;
      end = e->total_num + 1;
    } else {
      start = 1;
      end = e->args.num_agents + 1;
    }
    for (i = start; i < end; ++i) {
      if (unit_exists(e, i)) {
        order[n].id = i;
        order[n].d_norm = distance(e, agent_id, i, 1);
        order[n].d_raw = distance(e, agent_id, i, 0);
        ++n;
      }
    }
  }
  for (i = 1; i < n; ++i) { /* stable insertion sort == list.sort(key=d_norm) */
    near_t t = order[i];
    for (j = i - 1; j >= 0 && order[j].d_norm > t.d_norm; --j) order[j + 1] = order[j];
    order[j + 1] = t;
  }
  return n;
}

/* env_base.py:464-487 */
static int correct_angle_sign(const unit_t* opp_unit, const unit_t* ag_unit) {
  double x = opp_unit->lon, y = opp_unit->lat, a = opp_unit->heading;
  double rad = pymod(a, 360) * (M_PI / 180.0); /* math.radians */
  double x1 = x + pyround3(sin(rad));
  double y1 = y + pyround3(cos(rad));
  double xc = ag_unit->lon, yc = ag_unit->lat;
  double val = (x1 - x) * (yc - y) - (xc - x) * (y1 - y);
  return val < 0 ? 1 : -1;
}

/* ------------------------------------------------------------------ observations */
static int shot_flag(const unit_t* u) {
  int shot = u->cannon_current_burst_secs > 0;
  if (u->ac_type == 1) shot = shot || u->actual_missile != 0;
  return shot;
}
static double hdg_feature(const unit_t* u) { return clip(pymod(u->heading, 359) / 359, 0, 1); }

/* env_base.py:166-183 */
static int friendly_ac_values(orc_env_t* e, int agent_id, int fri_id, double* st) {
  int k;
  if (!fri_id || !unit_exists(e, fri_id)) {
    for (k = 0; k < 5; ++k) st[k] = 0;
    return 5;
  }
  relative_position(e, U(e, fri_id)->lat, U(e, fri_id)->lon, &st[0], &st[1]);
  st[2] = focus_angle(e, agent_id, fri_id, 1);
  st[3] = focus_angle(e, fri_id, agent_id, 1);
  st[4] = distance(e, agent_id, fri_id, 1);
  return 5;
}

/* env_base.py:185-212; mode 0 "fight", 1 "esc" (the "HighLevel" branch lives in env_hier) */
static int opp_ac_values(orc_env_t* e, int mode, int opp_id, int agent_id, double dist, double* st) {
  const unit_t* unit = U(e, opp_id);
  int n = 0;
  relative_position(e, unit->lat, unit->lon, &st[0], &st[1]);
  n = 2;
  st[n++] = clip(unit->speed / unit->max_speed, 0, 1);
  st[n++] = hdg_feature(unit);
  st[n++] = heading_diff(e, opp_id, agent_id);
  if (mode == 0) {
    st[n++] = focus_angle(e, opp_id, agent_id, 1);
    st[n++] = aspect_angle(e, agent_id, opp_id);
  } else {
    st[n++] = focus_angle(e, agent_id, opp_id, 1);
    st[n++] = focus_angle(e, opp_id, agent_id, 1);
  }
  st[n++] = dist;
  st[n++] = shot_flag(unit);
  return n;
}

/* env_base.py:111-135 */
static int fight_state_values(orc_env_t* e, int agent_id, const near_t* opp, int fri_id, double* st) {
  const unit_t* unit = U(e, agent_id);
  int n = 0;
  relative_position(e, unit->lat, unit->lon, &st[0], &st[1]);
  n = 2;
  st[n++] = clip(unit->speed / unit->max_speed, 0, 1);
  st[n++] = hdg_feature(unit);
  st[n++] = focus_angle(e, agent_id, opp->id, 1);
  st[n++] = aspect_angle(e, opp->id, agent_id);
  st[n++] = heading_diff(e, agent_id, opp->id);
  st[n++] = opp->d_norm;
  st[n++] = clip(unit->cannon_remain_secs / unit->cannon_max, 0, 1);
  if (unit->ac_type == 1) {
    st[n++] = clip((double)unit->missile_remain / unit->rocket_max, 0, 1);
    st[n++] = e->missile_wait[agent_id] == 0;
    st[n++] = (unit->actual_missile != 0) || (unit->cannon_current_burst_secs > 0);
  } else {
    st[n++] = unit->cannon_current_burst_secs > 0;
  }
  n += opp_ac_values(e, 0, opp->id, agent_id, opp->d_norm, st + n);
  n += friendly_ac_values(e, agent_id, fri_id, st + n);
  return n;
}

/* env_base.py:137-164 */
static int esc_state_values(orc_env_t* e, int agent_id, const near_t* opps, int n_opps, int fri_id,
                            double* st) {
  const unit_t* unit = U(e, agent_id);
  int n = 0, k, filled = 0;
  relative_position(e, unit->lat, unit->lon, &st[0], &st[1]);
  n = 2;
  st[n++] = clip(unit->speed / unit->max_speed, 0, 1);
  st[n++] = hdg_feature(unit);
  st[n++] = clip(unit->cannon_remain_secs / unit->cannon_max, 0, 1);
  if (unit->ac_type == 1) st[n++] = clip((double)unit->missile_remain / unit->rocket_max, 0, 1);
  st[n++] = shot_flag(unit);
  for (k = 0; k < n_opps; ++k) {
    filled += opp_ac_values(e, 1, opps[k].id, agent_id, opps[k].d_norm, st + n + filled);
    if (filled == 18) break;
  }
  for (; filled < 18; ++filled) st[n + filled] = 0;
  n += 18;
  n += friendly_ac_values(e, agent_id, fri_id, st + n);
  return n;
}

int orc_obs_len(const orc_args_t* args, int agent_id) {
  int ac1 = (agent_id == 1 || agent_id == 3);
  if (args->agent_mode == 0) return ac1 ? OBS_AC1 : OBS_AC2;
  return ac1 ? OBS_ESC_AC1 : OBS_ESC_AC2;
}
static int obs_len_mode(int mode, int agent_id) {
  int ac1 = (agent_id == 1 || agent_id == 3);
  if (mode == 0) return ac1 ? OBS_AC1 : OBS_AC2;
  return ac1 ? OBS_ESC_AC1 : OBS_ESC_AC2;
}

/* env_hetero.py:65-103 for ONE ag_id; writes float32 obs, updates opp_to_attack */
static void lowlevel_state_one(orc_env_t* e, int mode, int ag_id, float* out) {
  int len = obs_len_mode(mode, ag_id), k;
  double st[64];
  near_t opps[ORC_MAX_AC];
  int fri = ag_id <= e->args.num_agents ? (ag_id == 2 ? 1 : 2) : (ag_id == 4 ? 3 : 4);
  e->opp_to_attack[ag_id] = 0;
  for (k = 0; k < len; ++k) out[k] = 0.0f;
  if (unit_exists(e, ag_id)) {
    int n_opps = nearby_object(e, ag_id, 0, opps);
    if (n_opps) {
      int n = mode == 0 ? fight_state_values(e, ag_id, &opps[0], fri, st)
                        : esc_state_values(e, ag_id, opps, n_opps, fri, st);
      e->opp_to_attack[ag_id] = opps[0].id;
      if (n != len) e->error |= 16; /* the reference asserts (env_hetero.py:96) */
      for (k = 0; k < len; ++k) out[k] = (float)st[k];
    }
  }
}

/* env_hetero.py:62-63 */
static void state(orc_env_t* e, float* obs1, float* obs2) {
  lowlevel_state_one(e, e->args.agent_mode, 1, obs1);
  lowlevel_state_one(e, e->args.agent_mode, 2, obs2);
}

/* ------------------------------------------------------------------ actions */
/* env_base.py:214-238 (mode "LowLevel") */
static void take_base_action_mode(orc_env_t* e, int highlevel, unit_t* unit, int unit_id, int opp_id,
                                  const int32_t* act, double* rewards) {
  set_heading(e, unit, pymod(unit->heading + (act[0] - 6) * 15, 360));
  set_speed(e, unit, 100 + ((unit->max_speed - 100) / 8) * act[1]);
  if (act[2] != 0 && unit->cannon_remain_secs > 0) {
    fire_cannon(unit);
    if (!highlevel && unit_id <= e->args.num_agents)
      if (e->args.agent_mode == 1 && unit->cannon_remain_secs < 90) rewards[unit_id] -= 0.1;
  }
  if (unit->ac_type == 1 && act[3] != 0) {
    if (opp_id && unit->missile_remain > 0 && !unit->actual_missile && e->missile_wait[unit_id] == 0) {
      fire_missile(e, unit, U(e, opp_id));
      e->missile_wait[unit_id] = highlevel ? orc_rng_randint(&e->rng_g, 8, 12) : orc_rng_randint(&e->rng_g, 7, 17);
      if (!highlevel && unit_id <= e->args.num_agents)
        if (e->args.agent_mode == 1 && unit->missile_remain < 3) rewards[unit_id] -= 0.1;
    }
  }
  if (e->missile_wait[unit_id] > 0 && !unit->actual_missile)
    e->missile_wait[unit_id] = e->missile_wait[unit_id] - 1;
}

static void take_base_action(orc_env_t* e, unit_t* unit, int unit_id, int opp_id, const int32_t* act,
                             double* rewards) {
  take_base_action_mode(e, 0, unit, unit_id, opp_id, act, rewards);
}

/* env_hetero.py:118-123 */
static void opp_level1(orc_env_t* e, unit_t* unit, int unit_id) {
  if (!unit->actual_missile && (e->steps % 40) < 3 && orc_rng_randint(&e->rng_g, 0, 1) != 0 &&
      e->missile_wait[unit_id] == 0 && unit->ac_type == 1) {
    near_t d_ag[ORC_MAX_AC];
    if (nearby_object(e, unit_id, 0, d_ag)) {
      fire_missile(e, unit, U(e, d_ag[0].id));
      e->missile_wait[unit_id] = 5;
    }
  }
}

/* env_hetero.py:125-136 */
static void opp_level2(orc_env_t* e, unit_t* unit, int unit_id) {
  fire_cannon(unit);
  if (e->steps <= 5 || (e->steps % orc_rng_randint(&e->rng_g, 35, 45)) <= 5) {
    int r = orc_rng_randint(&e->rng_g, 0, 1);
    double s;
    set_heading(e, unit, pymod(unit->heading + (r ? -1 : 1) * 90, 360));
    s = 100 + orc_rng_randint(&e->rng_g, 0, 4) * 75;
    set_speed(e, unit, s);
  }
  opp_level1(e, unit, unit_id); /* identical missile rule, env_hetero.py:132-136 */
}

/* env_hetero.py:227-245 */
static void escaping_opp(orc_env_t* e, const unit_t* unit, double* heading, double* speed, int* fire) {
  double y, x;
  relative_position(e, unit->lat, unit->lon, &y, &x);
  if (y < 0.5) {
    if (x < 0.5)
      *heading = (int)orc_rng_uniform(&e->rng_g, 30, 60);
    else
      *heading = (int)orc_rng_uniform(&e->rng_g, 300, 330);
  } else {
    if (x < 0.5)
      *heading = (int)orc_rng_uniform(&e->rng_g, 120, 150);
    else
      *heading = (int)orc_rng_uniform(&e->rng_g, 210, 240);
  }
  *speed = (int)orc_rng_uniform(&e->rng_g, 300, 600);
  *fire = orc_rng_randint(&e->rng_g, 0, 1) != 0;
}

/* env_hetero.py:247-271 */
static void hardcoded_opp(orc_env_t* e, const unit_t* opp_unit, int opp_id, int* opp, double* heading,
                          double* speed, int* fire, int* fire_m) {
  near_t d_agt[ORC_MAX_AC];
  int n = nearby_object(e, opp_id, 0, d_agt);
  *heading = opp_unit->heading;
  *fire = 0;
  *fire_m = 0;
  *opp = 0;
  *speed = (int)orc_rng_uniform(&e->rng_g, 100, 400);
  if (n) {
    int sign = correct_angle_sign(opp_unit, U(e, d_agt[0].id));
    double r = orc_rng_uniform(&e->rng_g, 0.7, 1.3);
    double focus = focus_angle(e, opp_id, d_agt[0].id, 0);
    if (d_agt[0].d_norm > 0.008 && focus > 4) *heading = pymod(*heading + r * sign * focus, 360);
    if (d_agt[0].d_norm > 0.05)
      *speed = focus < 30 ? (int)orc_rng_uniform(&e->rng_g, 500, 800)
                          : (int)orc_rng_uniform(&e->rng_g, 100, 500);
    *fire = d_agt[0].d_norm < 0.03 && focus < 10;
    *fire_m = d_agt[0].d_norm < 0.09 && focus < 5;
    *opp = d_agt[0].id;
  }
  if (opp_unit->ac_type == 2) *speed = clip(*speed, 0, 600);
}

/* env_hetero.py:138-158 */
static void opp_level3(orc_env_t* e, unit_t* unit, int unit_id) {
  int opp = 0, fire = 0, fire_m = 0;
  double heading, speed;
  if (e->steps % 60 == 0 && !e->hardcoded_opps_escaping) {
    e->hardcoded_opps_escaping = orc_rng_randint(&e->rng_g, 0, 1) != 0;
    if (e->hardcoded_opps_escaping) e->opps_escaping_time = (int)orc_rng_uniform(&e->rng_g, 20, 30);
  }
  if (e->hardcoded_opps_escaping) {
    escaping_opp(e, unit, &heading, &speed, &fire);
    e->opps_escaping_time -= 1;
    if (e->opps_escaping_time <= 0) e->hardcoded_opps_escaping = 0;
  } else {
    hardcoded_opp(e, unit, unit_id, &opp, &heading, &speed, &fire, &fire_m);
  }
  set_heading(e, unit, heading);
  set_speed(e, unit, speed);
  if (fire) fire_cannon(unit);
  if (fire_m && opp && !unit->actual_missile && e->missile_wait[unit_id] == 0 && unit->ac_type == 1) {
    fire_missile(e, unit, U(e, opp));
    e->missile_wait[unit_id] = 10;
  }
}

/* env_base.py:349-398 */
static void policy_actions(orc_env_t* e, int policy_type, int agent_id, unit_t* unit, int32_t* act) {
  float obs[64];
  int len = obs_len_mode(policy_type, agent_id);
  lowlevel_state_one(e, policy_type, agent_id, obs);
  act[0] = act[1] = act[2] = act[3] = 0;
  if (e->policy_fn)
    e->policy_fn(e->policy_user, agent_id, unit->ac_type, policy_type, e->policy_set, obs, len, act);
  else
    e->error |= 32;
}

/* env_base.py:240-310 (mode "LowLevel") */
static void combat_rewards(orc_env_t* e, const event_t* events, int n_events,
                           double opp_stats[][2], double rews[], int destroyed[]) {
  double s = e->args.rew_scale;
  int i, k, na = e->args.num_agents;
  for (i = 1; i <= e->total_num; ++i) {
    if (unit_exists(e, i)) {
      unit_t* u = U(e, i);
      if (!in_boundary(e, u->lat, u->lon)) {
        remove_unit(e, i);
        if (i <= na) {
          rews[i] += -5 * s;
          destroyed[i] = 1;
          e->alive_agents -= 1;
        } else {
          e->alive_opps -= 1;
        }
      }
    }
  }
  for (k = 0; k < n_events; ++k) {
    const event_t* ev = &events[k];
    const unit_t* killer = U(e, ev->killer);
    if (ev->killer <= na) {
      if (ev->destroyed >= na + 1 && ev->destroyed <= e->total_num) {
        if (e->args.agent_mode == 0) {
          if (ev->origin >= e->total_num + 1)
            rews[ev->killer] +=
                shifted_range((double)killer->missile_remain / killer->rocket_max, 0, 1, 1, 1.5) * s;
          else
            rews[ev->killer] += (shifted_range(killer->cannon_remain_secs / killer->cannon_max, 0, 1, 0.5, 1) +
                                 shifted_range(opp_stats[ev->killer][0], 0, 1, 0.5, 1)) * s;
        }
        e->alive_opps -= 1;
      } else if (ev->destroyed <= na) {
        rews[ev->killer] += -2 * s;
        if (e->args.friendly_punish) {
          rews[ev->destroyed] += -2 * s;
          destroyed[ev->destroyed] = 1;
        }
        e->alive_agents -= 1;
      }
    } else if (ev->killer >= na + 1 && ev->killer <= e->total_num) {
      if (ev->destroyed <= na) {
        rews[ev->destroyed] += -2 * s;
        destroyed[ev->destroyed] = 1;
        e->alive_agents -= 1;
      } else if (ev->destroyed >= na + 1 && ev->destroyed <= e->total_num) {
        e->alive_opps -= 1;
      }
    }
  }
}

/* env_hetero.py:105-186 + 188-225 */
static void take_action(orc_env_t* e, const int32_t* actions, double* rewards, int32_t* present) {
  double opp_stats[ORC_MAX_AC + 1][2];
  double rews[ORC_MAX_AC + 1];
  int destroyed[ORC_MAX_AC + 1];
  event_t events[ORC_MAX_UNITS];
  int i, n_events, na = e->args.num_agents;
  e->steps += 1;
  memset(opp_stats, 0, sizeof opp_stats);
  memset(rews, 0, sizeof rews);
  memset(destroyed, 0, sizeof destroyed);
  for (i = 1; i <= na; ++i) {
    rewards[i] = 0;
    present[i] = 0;
  }
  for (i = 1; i <= e->total_num; ++i) {
    if (unit_exists(e, i)) {
      unit_t* u = U(e, i);
      if (i <= na || e->args.level >= 4) {
        int32_t pol_act[4];
        const int32_t* act;
        if (i >= na + 1) {
          policy_actions(e, e->opp_mode, i, u, pol_act);
          act = pol_act;
        } else {
          act = actions + 4 * (i - 1);
          rewards[i] = 0;
          present[i] = 1;
          if (unit_exists(e, e->opp_to_attack[i])) {
            opp_stats[i][0] = focus_angle(e, e->opp_to_attack[i], i, 1);
            opp_stats[i][1] = distance(e, i, e->opp_to_attack[i], 0);
          }
        }
        take_base_action(e, u, i, e->opp_to_attack[i], act, rewards);
      } else {
        if (e->args.level == 1)
          opp_level1(e, u, i);
        else if (e->args.level == 2)
          opp_level2(e, u, i);
        else if (e->args.level == 3)
          opp_level3(e, u, i);
      }
    }
  }
  n_events = do_tick(e, events);
  /* _get_rewards, env_hetero.py:188-225 */
  combat_rewards(e, events, n_events, opp_stats, rews, destroyed);
  if (e->args.agent_mode == 1 && e->args.esc_dist_rew) {
    for (i = 1; i <= na; ++i) {
      if (unit_exists(e, i)) {
        near_t opps[ORC_MAX_AC];
        int n = nearby_object(e, i, 0, opps), j;
        for (j = 1; j <= n; ++j) {
          const near_t* o = &opps[j - 1];
          if (o->d_raw < 0.06) {
            rews[i] += -0.02 / j;
            if (U(e, i)->speed < 200) rews[i] += -0.02 / j;
          } else if (o->d_raw > 0.13) {
            rews[i] += 0.02 / j;
            if (U(e, i)->speed > 500) rews[i] += 0.02 / j;
          }
        }
      }
    }
  }
  for (i = 1; i <= na; ++i) {
    if (unit_exists(e, i) || destroyed[i]) {
      if (!present[i]) e->error |= 64; /* KeyError in the reference */
      if (e->args.glob_frac > 0 && e->args.agent_mode == 0)
        rewards[i] += rews[i] + e->args.glob_frac * rews[i % 2 + 1];
      else
        rewards[i] += rews[i];
    }
  }
}

/* ------------------------------------------------------------------ reset */
/* env_base.py:489-549 */
static void sample_state(orc_env_t* e, int group, int i, int r, double* x, double* y, int* a) {
  orc_rng_t* g = &e->rng_g;
  int level = e->args.level;
  *x = 0;
  *y = 0;
  *a = 0;
  if (group == 0) {
    if (level == 1) {
      if (r == 1) {
        *x = orc_rng_uniform(g, 7.12, 7.14);
        *y = orc_rng_uniform(g, 5.1 + i * 0.1, 5.11 + i * 0.1);
        *a = orc_rng_randint(g, 30, 150);
      } else {
        *x = orc_rng_uniform(g, 7.16, 7.17);
        *y = orc_rng_uniform(g, 5.1 + i * 0.1, 5.11 + i * 0.1);
        *a = orc_rng_randint(g, 200, 330);
      }
    } else if (level == 2) {
      if (r == 1) {
        *x = orc_rng_uniform(g, 7.08, 7.13);
        *y = orc_rng_uniform(g, 5.08 + i * 0.1, 5.13 + i * 0.1);
        *a = orc_rng_randint(g, 0, 180);
      } else {
        *x = orc_rng_uniform(g, 7.18, 7.23);
        *y = orc_rng_uniform(g, 5.08 + i * 0.1, 5.13 + i * 0.1);
        *a = orc_rng_randint(g, 180, 359);
      }
    } else {
      if (r == 1) {
        *x = orc_rng_uniform(g, 7.07, 7.12);
        *y = orc_rng_uniform(g, 5.09 + i * 0.1, 5.12 + i * 0.1);
        *a = orc_rng_randint(g, 0, 270);
      } else {
        *x = orc_rng_uniform(g, 7.18, 7.23);
        *y = orc_rng_uniform(g, 5.09 + i * 0.1, 5.12 + i * 0.1);
        *a = orc_rng_randint(g, 90, 359);
      }
    }
  } else {
    if (level == 1) {
      if (r == 1) {
        *x = orc_rng_uniform(g, 7.16, 7.17);
        *y = orc_rng_uniform(g, 5.1 + i * 0.1, 5.11 + i * 0.1);
      } else {
        *x = orc_rng_uniform(g, 7.12, 7.14);
        *y = orc_rng_uniform(g, 5.1 + i * 0.1, 5.11 + i * 0.1);
      }
    } else if (level == 2) {
      if (r == 1) {
        *x = orc_rng_uniform(g, 7.18, 7.23);
        *y = orc_rng_uniform(g, 5.08 + i * 0.1, 5.13 + i * 0.1);
        *a = orc_rng_randint(g, 0, 359);
      } else {
        *x = orc_rng_uniform(g, 7.08, 7.13);
        *y = orc_rng_uniform(g, 5.08 + i * 0.1, 5.13 + i * 0.1);
        *a = orc_rng_randint(g, 0, 359);
      }
    } else {
      if (r == 1) {
        *x = orc_rng_uniform(g, 7.18, 7.23);
        *y = orc_rng_uniform(g, 5.09 + i * 0.1, 5.12 + i * 0.1);
        *a = orc_rng_randint(g, 0, 359);
      } else {
        *x = orc_rng_uniform(g, 7.07, 7.12);
        *y = orc_rng_uniform(g, 5.09 + i * 0.1, 5.12 + i * 0.1);
        *a = orc_rng_randint(g, 0, 359);
      }
    }
  }
}

/* env_base.py:551-585 (mode "LowLevel") */
static void reset_scenario(orc_env_t* e) {
  int r = orc_rng_randint(&e->rng_g, 1, 2);
  int group, i;
  for (group = 0; group < 2; ++group) {
    int count = group == 0 ? e->args.num_agents : e->args.num_opps;
    for (i = 0; i < count; ++i) {
      double x, y;
      int a, ac;
      unit_t u;
      sample_state(e, group, i, r, &x, &y, &a);
      ac = i <= 1 ? i + 1 : orc_rng_randint(&e->rng_g, 1, 2);
      memset(&u, 0, sizeof u);
      u.kind = ac == 1 ? KIND_AC1 : KIND_AC2;
      u.lat = y;
      u.lon = x;
      u.heading = a;
      if (u.heading >= 360 || u.heading < 0) e->error |= 1;
      u.speed = (e->args.level <= 2 && group == 1) ? 0 : 100;
      u.new_heading = u.heading;
      u.new_speed = u.speed;
      u.max_speed = ac == 1 ? 900 : 600;
      u.cannon_remain_secs = 200;
      u.cannon_max = 200;
      u.missile_remain = ac == 1 ? 5 : 0;
      u.rocket_max = ac == 1 ? 5 : 0;
      u.friendly_check = e->args.friendly_kill;
      u.group = group;
      u.ac_type = ac;
      if (e->args.level <= 4 && group == 1) {
        u.cannon_max = u.cannon_remain_secs = 400;
        if (ac == 1) u.missile_remain = u.rocket_max = 8;
      } else if (e->args.level == 5) {
        u.cannon_max = u.cannon_remain_secs = 300;
        if (ac == 1) u.missile_remain = u.rocket_max = 6;
      }
      add_unit(e, &u);
      if (group == 0)
        e->alive_agents += 1;
      else
        e->alive_opps += 1;
    }
  }
}


/* ================================================================== HighLevelEnv (envs/env_hier.py) */
/* opp_ac_values(mode="HighLevel"), env_base.py:185-212: 10 values */
static int opp_ac_values_hl(orc_env_t* e, int opp_id, int agent_id, double dist, double* st) {
  const unit_t* unit = U(e, opp_id);
  int n = 0;
  relative_position(e, unit->lat, unit->lon, &st[0], &st[1]);
  n = 2;
  st[n++] = clip(unit->speed / unit->max_speed, 0, 1);
  st[n++] = hdg_feature(unit);
  st[n++] = heading_diff(e, opp_id, agent_id);
  st[n++] = focus_angle(e, agent_id, opp_id, 1);
  st[n++] = focus_angle(e, opp_id, agent_id, 1);
  st[n++] = aspect_angle(e, agent_id, opp_id);
  st[n++] = aspect_angle(e, opp_id, agent_id);
  st[n++] = dist;
  return n;
}

/* HighLevelEnv.state, env_hier.py:49-98 */
static void hier_state(orc_env_t* e, float* obs) {
  int ag, k, na = e->args.num_agents;
  for (ag = 1; ag <= e->total_num; ++ag) {
    e->ota_n[ag] = 0;
    if (ag <= na) {
      float* out = obs + (size_t)(ag - 1) * ORC_OBS_HL;
      for (k = 0; k < ORC_OBS_HL; ++k) out[k] = 0.0f;
      if (unit_exists(e, ag)) {
        near_t opps[ORC_MAX_AC], fri[ORC_MAX_AC];
        int n_opps = nearby_object(e, ag, 0, opps);
        if (n_opps) {
          double st[64];
          const unit_t* unit = U(e, ag);
          int n = 0, filled = 0, n_fri, ffilled = 0;
          relative_position(e, unit->lat, unit->lon, &st[0], &st[1]);
          n = 2;
          st[n++] = clip(unit->speed / unit->max_speed, 0, 1);
          st[n++] = hdg_feature(unit);
          for (k = 0; k < n_opps; ++k) {
            filled += opp_ac_values_hl(e, opps[k].id, ag, opps[k].d_norm, st + n + filled);
            e->ota_list[ag][e->ota_n[ag]].id = opps[k].id;
            e->ota_list[ag][e->ota_n[ag]].d_norm = opps[k].d_norm;
            e->ota_list[ag][e->ota_n[ag]].d_raw = opps[k].d_raw;
            e->ota_n[ag] += 1;
            if (filled == 20) break;
          }
          for (; filled < 20; ++filled) st[n + filled] = 0;
          n += 20;
          n_fri = nearby_object(e, ag, 1, fri);
          for (k = 0; k < n_fri; ++k) {
            ffilled += friendly_ac_values(e, ag, fri[k].id, st + n + ffilled);
            if (ffilled == 10) break;
          }
          for (; ffilled < 10; ++ffilled) st[n + ffilled] = 0;
          n += 10;
          if (n != ORC_OBS_HL) e->error |= 16;
          for (k = 0; k < ORC_OBS_HL; ++k) out[k] = (float)st[k];
        }
      }
    } else if (unit_exists(e, ag)) {
      near_t opps[ORC_MAX_AC];
      int n_opps = nearby_object(e, ag, 0, opps);
      for (k = 0; k < n_opps; ++k) {
        e->ota_list[ag][k].id = opps[k].id;
        e->ota_list[ag][k].d_norm = opps[k].d_norm;
        e->ota_list[ag][k].d_raw = opps[k].d_raw;
      }
      e->ota_n[ag] = n_opps;
    }
  }
}

/* python list indexing l[idx] with negative wrap; returns -1 when out of range (IndexError) */
static int py_index(int n, int idx) {
  if (idx < 0) idx += n;
  return (idx >= 0 && idx < n) ? idx : -1;
}

/* HighLevelEnv.lowlevel_state, env_hier.py:100-112 */
static int hier_lowlevel_state(orc_env_t* e, int mode, int agent_id, float* out) {
  double st[64];
  near_t fri[ORC_MAX_AC], opps[ORC_MAX_AC];
  unit_t* unit = U(e, agent_id);
  int n_fri = nearby_object(e, agent_id, 1, fri), k, n, len;
  int fri_id = n_fri ? fri[0].id : 0;
  int ca = e->commander_actions[agent_id];
  for (k = 0; k < e->ota_n[agent_id]; ++k) {
    opps[k].id = e->ota_list[agent_id][k].id;
    opps[k].d_norm = e->ota_list[agent_id][k].d_norm;
    opps[k].d_raw = e->ota_list[agent_id][k].d_raw;
  }
  if (mode == 0) {
    int idx = py_index(e->ota_n[agent_id], ca - 1);
    if (idx < 0) { e->error |= 128; idx = 0; }
    n = fight_state_values(e, agent_id, &opps[idx], fri_id, st);
    len = unit->ac_type == 1 ? OBS_AC1 : OBS_AC2;
  } else {
    n = esc_state_values(e, agent_id, opps, e->ota_n[agent_id], fri_id, st);
    len = unit->ac_type == 1 ? OBS_ESC_AC1 : OBS_ESC_AC2;
  }
  if (n != len) e->error |= 16;
  for (k = 0; k < len; ++k) out[k] = (float)st[k];
  return len;
}

/* HighLevelEnv._action_assess, env_hier.py:142-190 */
static void hier_action_assess(orc_env_t* e, double* rewards) {
  int i, na = e->args.num_agents;
  for (i = 1; i <= e->total_num; ++i) {
    if (unit_exists(e, i)) {
      if (i <= na) {
        rewards[i] = 0;
        if (e->commander_actions[i] > 0) {
          int idx = py_index(e->ota_n[i], e->commander_actions[i] - 1);
          int opp_id = 0;
          if (idx >= 0) {
            opp_id = e->ota_list[i][idx].id;
          } else {
            e->commander_actions[i] = 1;
          }
          if (!opp_id) rewards[i] = -0.1;
          if (e->args.hier_action_assess && opp_id) {
            if (distance(e, i, opp_id, 0) < 0.1 && focus_angle(e, i, opp_id, 0) < 15 && focus_angle(e, opp_id, i, 0) > 40)
              rewards[i] = 0.1;
            else
              rewards[i] = 0;
          }
        } else if (e->args.hier_action_assess) {
          if (e->ota_n[i] > 0) {
            int cl_opp = e->ota_list[i][0].id;
            if (distance(e, cl_opp, i, 0) < 0.1 && focus_angle(e, cl_opp, i, 0) < 15 && focus_angle(e, i, cl_opp, 0) > 40)
              rewards[i] = 0.1;
          } else {
            e->error |= 128; /* IndexError in the reference */
          }
        }
      } else {
        /* Fraction(ratio, 100).limit_denominator().as_integer_ratio() */
        int p = e->args.hier_opp_fight_ratio, q = 100, a = p, b = q, ag_id;
        while (b) { int t = a % b; a = b; b = t; }
        if (a > 0) { p /= a; q /= a; }
        if (orc_rng_choice2(&e->rng_g, (double)(q - p), (double)p)) {
          int possible = e->ota_n[i];
          if (possible > 1 && orc_rng_choice2(&e->rng_g, 1.0, 3.0))
            ag_id = orc_rng_randint(&e->rng_g, 2, possible);
          else
            ag_id = 1;
        } else {
          ag_id = 0;
        }
        e->commander_actions[i] = ag_id;
      }
    } else {
      if (i <= na) rewards[i] = 0;
      e->commander_actions[i] = -1;
    }
  }
}

/* HighLevelEnv._surrounding_event, env_hier.py:192-208 */
static int hier_surrounding_event(orc_env_t* e) {
  int i, j, event = 0, na = e->args.num_agents;
  for (i = 1; i <= na; ++i) {
    for (j = na + 1; j <= e->total_num; ++j) {
      if (unit_exists(e, i) && unit_exists(e, j)) {
        event = 0;
        if (distance(e, i, j, 0) < 0.1)
          if (focus_angle(e, i, j, 0) < 15 || focus_angle(e, j, i, 0) < 15) event = 1;
      }
      if (event) break;
    }
    if (event) break;
  }
  return event;
}

/* _combat_rewards(mode="HighLevel") (env_base.py:240-310) + HighLevelEnv._get_rewards (env_hier.py:210-224) */
static int hier_get_rewards(orc_env_t* e, double* rewards, const event_t* events, int n_events) {
  double s = e->args.rew_scale, rews[ORC_MAX_AC + 1];
  int destroyed[ORC_MAX_AC + 1], kill_event = 0, i, k, na = e->args.num_agents;
  memset(rews, 0, sizeof rews);
  memset(destroyed, 0, sizeof destroyed);
  for (i = 1; i <= e->total_num; ++i) {
    if (unit_exists(e, i)) {
      unit_t* u = U(e, i);
      if (!in_boundary(e, u->lat, u->lon)) {
        remove_unit(e, i);
        kill_event = 1;
        if (i <= na) {
          rews[i] += -2 * s;
          destroyed[i] = 1;
          e->alive_agents -= 1;
        } else {
          e->alive_opps -= 1;
        }
      }
    }
  }
  for (k = 0; k < n_events; ++k) {
    const event_t* ev = &events[k];
    if (ev->killer <= na) {
      if (ev->destroyed >= na + 1 && ev->destroyed <= e->total_num) {
        rews[ev->killer] += 1; /* constant, unscaled (env_base.py:285) */
        e->alive_opps -= 1;
      } else if (ev->destroyed <= na) {
        e->alive_agents -= 1;
      }
    } else if (ev->killer >= na + 1 && ev->killer <= e->total_num) {
      if (ev->destroyed <= na) {
        rews[ev->destroyed] += -1 * s;
        destroyed[ev->destroyed] = 1;
        e->alive_agents -= 1;
      } else if (ev->destroyed >= na + 1 && ev->destroyed <= e->total_num) {
        e->alive_opps -= 1;
      }
    }
    kill_event = 1;
  }
  for (i = 1; i <= na; ++i) {
    if (unit_exists(e, i) || destroyed[i]) {
      if (e->args.glob_frac > 0) {
        double others = 0;
        int j;
        for (j = 1; j <= na; ++j)
          if (j != i) others += rews[j];
        rewards[i] += rews[i] + e->args.glob_frac * others;
      } else {
        rewards[i] += rews[i];
      }
    }
  }
  return kill_event;
}

/* HighLevelEnv._sample_state (env_hier.py:226-250) + _reset_scenario(mode="HighLevel") (env_base.py:551-585) */
static void hier_reset_scenario(orc_env_t* e) {
  int r = orc_rng_randint(&e->rng_g, 1, 2);
  int group, i;
  for (group = 0; group < 2; ++group) {
    int count = group == 0 ? e->args.num_agents : e->args.num_opps;
    for (i = 0; i < count; ++i) {
      double x, y, step = 0.4 / count;
      int a, ac, west = (group == 0) == (r == 1);
      unit_t u;
      x = west ? orc_rng_uniform(&e->rng_g, 7.07, 7.22) : orc_rng_uniform(&e->rng_g, 7.28, 7.43);
      y = orc_rng_uniform(&e->rng_g, 5.07 + i * step, 5.12 + i * step);
      a = orc_rng_randint(&e->rng_g, 0, 359);
      ac = i <= 1 ? i + 1 : orc_rng_randint(&e->rng_g, 1, 2);
      memset(&u, 0, sizeof u);
      u.kind = ac == 1 ? KIND_AC1 : KIND_AC2;
      u.lat = y;
      u.lon = x;
      u.heading = a;
      u.speed = (e->args.level <= 2 && group == 1) ? 0 : 100; /* args.level keeps its default 1 in train_hier.py */
      u.new_heading = u.heading;
      u.new_speed = u.speed;
      u.max_speed = ac == 1 ? 900 : 600;
      u.cannon_remain_secs = u.cannon_max = 300; /* env_base.py:575-578 */
      u.missile_remain = u.rocket_max = ac == 1 ? 8 : 0;
      u.friendly_check = e->args.friendly_kill;
      u.group = group;
      u.ac_type = ac;
      add_unit(e, &u);
      if (group == 0)
        e->alive_agents += 1;
      else
        e->alive_opps += 1;
    }
  }
}

void orc_hier_reset(orc_env_t* e, float* obs) {
  int i;
  e->steps = 0;
  e->alive_agents = 0;
  e->alive_opps = 0;
  e->hardcoded_opps_escaping = 0;
  e->opps_escaping_time = 0;
  for (i = 0; i <= ORC_MAX_AC; ++i) {
    e->missile_wait[i] = 0;
    e->opp_to_attack[i] = 0;
    e->ota_n[i] = 0;
    e->commander_actions[i] = -1;
  }
  memset(e->units, 0, sizeof e->units);
  e->next_unit_id = 1;
  e->utc_time = 0;
  hier_reset_scenario(e);
  hier_state(e, obs);
}

/* HHMARLBaseEnv.step + HighLevelEnv._take_action, env_base.py:79-109, env_hier.py:114-140 */
int orc_hier_step(orc_env_t* e, const int32_t* commander_actions, float* obs, double* rew, int32_t* info) {
  double rewards[ORC_MAX_AC + 1];
  event_t events[ORC_MAX_UNITS];
  int s = 0, kill_event = 0, situation_event = 0, i, na = e->args.num_agents;
  const int n_sub_steps = 15, min_sub_steps = 10;
  memset(rewards, 0, sizeof rewards);
  for (i = 1; i <= e->total_num; ++i) e->commander_actions[i] = i <= na ? commander_actions[i - 1] : -1;
  hier_action_assess(e, rewards);
  if (info)
    for (i = 1; i <= e->total_num; ++i) info[i] = e->commander_actions[i];
  while (s <= n_sub_steps && !kill_event && !situation_event) {
    int n_events;
    for (i = 1; i <= e->total_num; ++i) {
      if (unit_exists(e, i)) {
        unit_t* u = U(e, i);
        float pobs[64];
        int32_t act[4] = {0, 0, 0, 0};
        int mode = e->commander_actions[i] == 0 ? 1 : 0;
        int len = hier_lowlevel_state(e, mode, i, pobs);
        int idx = py_index(e->ota_n[i], e->commander_actions[i] - 1);
        if (e->policy_fn)
          e->policy_fn(e->policy_user, i, u->ac_type, mode, 0, pobs, len, act);
        else
          e->error |= 32;
        if (idx < 0) { e->error |= 128; idx = 0; }
        take_base_action_mode(e, 1, u, i, e->ota_list[i][idx].id, act, rewards);
      }
    }
    n_events = do_tick(e, events);
    kill_event = hier_get_rewards(e, rewards, events, n_events);
    if (s > min_sub_steps) situation_event = hier_surrounding_event(e);
    s += 1;
    e->steps += 1;
  }
  if (info) info[0] = s;
  for (i = 1; i <= na; ++i) rew[i - 1] = rewards[i];
  hier_state(e, obs);
  return e->alive_agents <= 0 || e->alive_opps <= 0 || e->steps >= e->args.horizon;
}

/* the `info` dict of HHMARLBaseEnv.step when args.eval_info is set (env_base.py:91-107; consumed by
 * evaluation.py:59-60, 66-82), evaluated on the env as orc_hier_step left it: units that exist NOW, the commander
 * action dict as _action_assess left it (env_hier.py:142-190).  out[12] = agents_win, opps_win, draw, agent_fight,
 * agent_escape, opp_fight, opp_escape, agent_steps, opp_steps, opp1, opp2, opp3. */
void orc_hier_eval_info(const orc_env_t* e, int32_t* out) {
  int i, na = e->args.num_agents;
  const int before_horizon = e->steps < e->args.horizon;
  memset(out, 0, 12 * sizeof(int32_t));
  out[0] = e->alive_opps <= 0 && before_horizon;
  out[1] = e->alive_agents <= 0 && before_horizon;
  out[2] = !before_horizon && e->alive_agents > 0 && e->alive_opps > 0;
  for (i = 1; i <= e->total_num; ++i) {
    const int v = e->commander_actions[i];
    if (!unit_exists(e, i)) continue;
    if (v > 0) { /* `if v:` -- None and 0 are falsy */
      if (i <= na) {
        out[3] += 1;
        out[7] += 1;
        if (v <= 3) out[8 + v] += 1;
      } else {
        out[5] += 1;
        out[8] += 1;
      }
    } else if (i <= na) {
      out[4] += 1;
      out[7] += 1;
    } else {
      out[6] += 1;
      out[8] += 1;
    }
  }
}

/* ------------------------------------------------------------------ public API */
orc_env_t* orc_env_create(const orc_args_t* args, uint64_t seed, uint32_t arena_id) {
  orc_env_t* e = (orc_env_t*)calloc(1, sizeof *e);
  if (!e) return 0;
  e->args = *args;
  e->total_num = args->num_agents + args->num_opps;
  e->rng_g.key[0] = e->rng_c.key[0] = (uint32_t)seed;
  e->rng_g.key[1] = e->rng_c.key[1] = (uint32_t)(seed >> 32);
  e->rng_g.arena = e->rng_c.arena = arena_id;
  e->rng_g.stream = 0;
  e->rng_c.stream = 1;
  e->next_unit_id = 1;
  return e;
}
void orc_env_destroy(orc_env_t* e) { free(e); }
void orc_env_set_policy_fn(orc_env_t* e, orc_policy_fn fn, void* user) {
  e->policy_fn = fn;
  e->policy_user = user;
}

/* env_base.py:62-77 + env_hetero.py:53-60 */
void orc_env_reset(orc_env_t* e, float* obs1, float* obs2) {
  int i;
  e->steps = 0;
  e->alive_agents = 0;
  e->alive_opps = 0;
  e->hardcoded_opps_escaping = 0;
  e->opps_escaping_time = 0;
  for (i = 0; i <= ORC_MAX_AC; ++i) {
    e->missile_wait[i] = 0;
    e->opp_to_attack[i] = 0;
  }
  memset(e->units, 0, sizeof e->units); /* new CmanoSimulator (env_base.py:75) */
  e->next_unit_id = 1;
  e->utc_time = 0;
  reset_scenario(e);
  e->opp_mode = 0;
  e->policy_set = 0;
  if (e->args.level == 5 && e->args.agent_mode == 0) {
    int k = orc_rng_randint(&e->rng_g, 3, 5);
    e->policy_set = k;
    e->opp_mode = k == 5 ? 1 : 0;
  }
  state(e, obs1, obs2);
}

/* env_base.py:79-109 */
int orc_env_step(orc_env_t* e, const int32_t* actions, float* obs1, float* obs2, double* rew,
                 int32_t* rew_present) {
  double rewards[ORC_MAX_AC + 1];
  int32_t present[ORC_MAX_AC + 1];
  int done, i;
  take_action(e, actions, rewards, present);
  done = e->alive_agents <= 0 || e->alive_opps <= 0 || e->steps >= e->args.horizon;
  state(e, obs1, obs2);
  for (i = 1; i <= e->args.num_agents; ++i) {
    rew[i - 1] = present[i] ? rewards[i] : 0.0;
    if (rew_present) rew_present[i - 1] = present[i];
  }
  return done;
}

void orc_env_get_state(const orc_env_t* e, orc_state_t* out) {
  int i;
  memset(out, 0, sizeof *out);
  for (i = 1; i <= e->total_num; ++i) {
    const unit_t* u = &e->units[i];
    int k = i - 1;
    out->lat[k] = u->lat;
    out->lon[k] = u->lon;
    out->heading[k] = u->heading;
    out->speed[k] = u->speed;
    out->new_heading[k] = u->new_heading;
    out->new_speed[k] = u->new_speed;
    out->cannon_remain[k] = u->cannon_remain_secs;
    out->cannon_burst[k] = u->cannon_current_burst_secs;
    out->cannon_max[k] = u->cannon_max;
    out->missile_remain[k] = u->missile_remain;
    out->rocket_max[k] = u->rocket_max;
    out->missile_wait[k] = e->missile_wait[i];
    out->alive[k] = u->active;
    out->has_missile[k] = u->actual_missile != 0;
    out->opp_to_attack[k] = e->opp_to_attack[i];
    out->ac_type[k] = u->ac_type;
    if (u->actual_missile) {
      const unit_t* m = &e->units[u->actual_missile];
      out->r_lat[k] = m->lat;
      out->r_lon[k] = m->lon;
      out->r_heading[k] = m->heading;
      out->r_new_heading[k] = m->new_heading;
      out->r_speed[k] = m->speed;
      out->r_alive[k] = m->active;
      out->r_target[k] = m->target;
      out->r_id[k] = m->id;
      out->r_age[k] = (int32_t)(e->utc_time - m->firing_time);
    }
  }
  out->steps = e->steps;
  out->alive_agents = e->alive_agents;
  out->alive_opps = e->alive_opps;
  out->escaping = e->hardcoded_opps_escaping;
  out->escaping_time = e->opps_escaping_time;
  out->next_unit_id = e->next_unit_id;
  out->opp_mode = e->opp_mode;
  out->policy_set = e->policy_set;
  out->error = e->error;
  out->draws_g = e->rng_g.draw;
  out->draws_c = e->rng_c.draw;
}

uint64_t orc_env_run_random(orc_env_t* e, uint64_t n_steps, uint64_t action_seed) {
  float obs1[32], obs2[32];
  double rew[2];
  int32_t act[8];
  uint64_t x = action_seed * 0x9E3779B97F4A7C15ull + 0x1234567ull, k, episodes = 0;
  static const int heads[4] = {13, 9, 2, 2};
  orc_env_reset(e, obs1, obs2);
  for (k = 0; k < n_steps; ++k) {
    int a, h;
    for (a = 0; a < 2; ++a)
      for (h = 0; h < 4; ++h) {
        x ^= x << 13;
        x ^= x >> 7;
        x ^= x << 17;
        act[a * 4 + h] = (int32_t)((x >> 33) % (uint64_t)heads[h]);
      }
    if (orc_env_step(e, act, obs1, obs2, rew, 0)) {
      orc_env_reset(e, obs1, obs2);
      ++episodes;
    }
  }
  return episodes;
}
