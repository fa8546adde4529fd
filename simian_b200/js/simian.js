/*
 * simian.js (simian-b200 facade) -- drop-in for SimianJS/simian.js + entity.js +
 * eventQ.js on the event-oriented service path, over the N-API addon
 * `simian_b200_napi` (simian_b200/js/napi_shim.cc), itself a 1:1 wrapper of the
 * C-ABI in include/simian_b200.h.
 *
 * Kept from the reference (same names / arity / error messages):
 *   new Simian(simName, startTime, endTime, minDelay, useMPI)   simian.js:36-84
 *   addEntity(name, entityClass, num, ...args)                  simian.js:211-229
 *   schedService(time, eventName, data, rx, rxId)               simian.js:170-190
 *   attachService(klass, name, fun)                             simian.js:207-209
 *   getEntity / getBaseRank / getOffsetRank / run / exit        simian.js:90-205
 *   class Entity { reqService, attachService }                  entity.js:21-72
 * Changed: handlers are CUDA device functions bound by name --
 *   `Node.prototype.generate = deviceFn("phold_generate")` -- and `data` is an
 *   integer (8-byte payload).  Coroutine processes are out of scope.
 *
 * NOT RUN in this repository's CI: the build image has no Node (SURVEY.md B).
 * tests/ cover the same calls through the Python mirror (simian_b200/simian.py).
 */
"use strict";
const native = require("./build/Release/simian_b200_napi.node");

class DeviceFn { constructor(name) { this.deviceFn = name; } }
const deviceFn = (name) => new DeviceFn(name);

function hash(str) {                     // SimianJS/hash.js:21-26
    var res = 5381, i = str.length;
    while (i--) res = 33 * res + str.charCodeAt(i);
    return res;
}

class Entity {
    constructor(name, out, engine, num) { // entity.js:21-29
        this.name = name; this.out = out; this.engine = engine; this.num = num;
    }
    toString() { return this.name + "(" + this.num + ")"; }
    reqService(offset, eventName, data, rx, rxId) { // entity.js:35-68 (host side: before run only)
        const engine = this.engine;
        if ((rx != null) && (offset < engine.minDelay)) {
            if (!engine.running) throw new Error("Cannot send an event when Simian is idle");
            throw new Error(this + " attempted to send with too little delay");
        }
        const time = engine.now + offset;
        if (time > engine.endTime) return;
        rx = (rx != null) ? rx : this.name;
        rxId = (rxId != null) ? rxId : this.num;
        engine._inject(time, eventName, data, rx, rxId, this.name, this.num);
    }
    attachService(name, fun) { this.engine.attachService(this.constructor, name, fun, this.name); }
}

class Simian {
    constructor(simName, startTime, endTime, minDelay, useMPI, rank, size) {
        this.name = simName; this.startTime = startTime;
        this.endTime = endTime || 1e100; this.minDelay = minDelay || 1;   // simian.js:41-42
        this.now = startTime; this.running = false;
        this.entities = {}; this.baseRanks = {}; this._classes = {};
        this.useMPI = !!useMPI;
        this.rank = useMPI ? rank : 0; this.size = useMPI ? size : 1;     // simian.js:71-79
        this._h = native.create(simName, startTime, endTime, minDelay, this.rank, this.size, -1);
        this._seq = 0;
    }
    getBaseRank(name) { return hash(name) % this.size; }                  // simian.js:192-194
    getOffsetRank(name, num) { return (this.baseRanks[name] + num) % this.size; }
    getEntity(name, num) { if (this.entities[name]) return this.entities[name][num]; }
    _register(name, klass) {
        if (this._classes[name]) return;
        this._classes[name] = klass;
        native.defineClass(this._h, name, klass.stateBytes || 0);
        // a subclass's own getBaseRank (simian.js:192-194) goes to the engine
        if (this.getBaseRank !== Simian.prototype.getBaseRank)
            native.setBaseRank(this._h, name, this.getBaseRank(name) % this.size);
        for (const key of Object.getOwnPropertyNames(klass.prototype)) {
            const f = klass.prototype[key];
            if (f instanceof DeviceFn) native.attachService(this._h, name, key, f.deviceFn);
        }
        if (klass.params) native.setClassParams(this._h, name, klass.params);
    }
    attachService(klass, name, fun, className) {                          // simian.js:207-209
        if (!(fun instanceof DeviceFn)) throw new Error("service handlers are device functions: use deviceFn(name)");
        klass.prototype[name] = fun;
        for (const c in this._classes)
            if (this._classes[c] === klass || c === className) native.attachService(this._h, c, name, fun.deviceFn);
    }
    addEntity(name, entityClass, num, ...theArgs) {                       // simian.js:211-229
        if (this.running) throw new Error("Cannot add new entity when Simian is running!");
        this._register(name, entityClass);
        if (!this.entities[name]) this.entities[name] = {};
        this.baseRanks[name] = this.getBaseRank(name);
        native.addEntities(this._h, name, num, 1);
        if (this.getOffsetRank(name, num) == this.rank)
            this.entities[name][num] = new entityClass(name, this.out, this, num, ...theArgs);
    }
    addEntities(name, entityClass, first, count) {  // bulk form of the addEntity loop
        if (this.running) throw new Error("Cannot add new entity when Simian is running!");
        this._register(name, entityClass);
        this.baseRanks[name] = this.getBaseRank(name);
        native.addEntities(this._h, name, first, count);
    }
    schedService(time, eventName, data, rx, rxId) {                       // simian.js:170-190
        if (time > this.endTime) return;
        native.schedService(this._h, time, eventName, data || 0, rx, rxId);
    }
    schedServiceBulk(time, eventName, data, rx, first, count, repeat) {
        native.schedServiceBulk(this._h, time, eventName, data || 0, rx, first, count, repeat || 1);
    }
    _inject(time, eventName, data, rx, rxId, tx, txId) {
        native.schedEvent(this._h, time, native.firstId(this._h, rx) + rxId,
            native.firstId(this._h, tx) + txId, native.internService(this._h, eventName),
            this._seq++, data || 0);
    }
    run() {                                                               // simian.js:99-168
        this.running = true;
        try { this.stats = native.run(this._h); }   // size > 1: see INTEGRATION.md (stepping calls + NCCL)
        finally { this.running = false; }
        this.now = this.stats.lastTime;
        console.log("SIMULATION FINISHED IN : " + this.stats.runSeconds + " (s)");
        console.log("SIMULATED EVENTS       : " + this.stats.totalEvents);
        console.log("MODEL-EVENTS/SECOND    : " + this.stats.totalEvents / this.stats.runSeconds);
        return this.stats;
    }
    exit() { if (this._h) { native.destroy(this._h); this._h = null; } } // simian.js:90-97
}

module.exports = { Simian, Entity, deviceFn, hash };
